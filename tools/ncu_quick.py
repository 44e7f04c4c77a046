#!/usr/bin/env python
"""Key numbers + stall reasons of one .ncu-rep (first kernel).  usage: tools/ncu_quick.py report.ncu-rep [topN source lines]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
def ncu_csv(*args):
    out = subprocess.run(["ncu", "-i", rep, "--csv"] + list(args), capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))
rows = ncu_csv("--page", "raw")
hdr, units, r = rows[0], rows[1], rows[2]
d = dict(zip(hdr, r))
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sass__inst_executed_shared_loads", "sass__inst_executed_global_loads",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_eligible.avg.per_cycle_active", "sm__maximum_warps_per_active_cycle_pct"]
print(d.get("Kernel Name", "")[:80])
for k in keys:
    if k in d: print("  %-62s %s %s" % (k, d[k], dict(zip(hdr, units)).get(k, "")))
st = [(k, float(d[k].replace(",", ""))) for k in hdr if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and d[k] not in ("", "n/a")]
tot = sum(v for _, v in st) or 1
print("  stalls:", ", ".join("%s %.0f%%" % (k.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot) for k, v in sorted(st, key=lambda x: -x[1])[:8]))
if topn:
    rows = ncu_csv("--page", "source", "--print-source", "cuda")
    # find header
    h = None
    for i, row in enumerate(rows):
        if "Source" in row and any("Instructions Executed" in c for c in row):
            h = i; break
    if h is not None:
        hd = rows[h]
        si = hd.index("Source"); ii = [j for j, c in enumerate(hd) if c == "Instructions Executed"][0]
        sj = [j for j, c in enumerate(hd) if c.startswith("Warp Stall Sampling (All")]
        fi = hd.index("File Path") if "File Path" in hd else None
        items = []
        for row in rows[h + 1:]:
            try: items.append((int(row[ii].replace(",", "") or 0), int(row[sj[0]].replace(",", "") or 0) if sj else 0, row[si].strip()[:110], row[0]))
            except Exception: pass
        ti = sum(x[0] for x in items) or 1; ts = sum(x[1] for x in items) or 1
        for x in sorted(items, key=lambda x: -x[1])[:topn]:
            print("  %5.1f%% inst %5.1f%% samp  L%-5s %s" % (100 * x[0] / ti, 100 * x[1] / ts, x[3], x[2]))
