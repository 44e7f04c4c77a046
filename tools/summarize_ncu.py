"""Turn an .ncu-rep (ncu --set full --import-source on) into the markdown summary kept under profiles/.

usage: python tools/summarize_ncu.py report.ncu-rep "title" "command that was profiled" > profiles/xyz.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__sass_branch_targets_threads_divergent.sum",
]


def ncu_csv(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv"] + list(args), capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = ncu_csv(rep, "--page", "raw")
    hdr, units = rows[0], rows[1]
    print("# %s\n" % title)
    print("Command: `%s`\n" % cmd)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("## %s\n" % d["Kernel Name"].split("(")[0])
        print("| metric | unit | value |\n|---|---|---|")
        for k in KEYS:
            if k in d:
                print("| %s | %s | %s |" % (k, u[k], d[k]))
        st = [(h, float(d[h].replace(",", ""))) for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and d[h] not in ("", "nan")]
        tot = sum(v for _, v in st) or 1.0
        print("\nWarp-state samples (share of all pc samples):\n")
        print("| stall reason | share |\n|---|---|")
        for h, v in sorted(st, key=lambda x: -x[1])[:9]:
            print("| %s | %.1f %% |" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot))
        print()
    # per-source-line instruction shares
    rows = ncu_csv(rep, "--page", "source", "--print-source", "cuda,sass")
    cur = None
    lines = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0].isdigit():
            try:
                lines.append((cur, int(r[0]), r[1].strip()[:100], int(r[7]), int(r[6])))
            except Exception:
                pass
    if lines:
        ti = sum(l[3] for l in lines) or 1
        ts = sum(l[4] for l in lines) or 1
        byf = collections.Counter()
        bys = collections.Counter()
        for l in lines:
            byf[l[0]] += l[3]
            bys[l[0]] += l[4]
        print("## Instruction and sample shares by source file (last kernel in the report)\n")
        print("| file | warp instructions | pc samples |\n|---|---|---|")
        for f, c in byf.most_common(12):
            print("| %s | %.1f %% | %.1f %% |" % (f, 100 * c / ti, 100 * bys[f] / ts))
        print("\n## Hottest source lines by executed warp instructions\n")
        print("| inst | samples | where | source |\n|---|---|---|---|")
        for l in sorted(lines, key=lambda x: -x[3])[:25]:
            print("| %.1f %% | %.1f %% | %s:%d | `%s` |" % (100 * l[3] / ti, 100 * l[4] / ts, l[0], l[1], l[2].replace("|", "\\|")))


if __name__ == "__main__":
    main()
