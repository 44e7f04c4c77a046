#!/usr/bin/env python
"""Instructions / stall samples of one kernel in an .ncu-rep (--import-source on) per source function and per line.

usage: tools/ncu_by_line.py report.ncu-rep [top_lines]
Function attribution: the nearest preceding line of the same file that looks like a function header (CB_DEV / CB_MEM /
template / __global__); inlined code counts for the function it was written in."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_re = re.compile(r"^\s*(template\s*<.*>\s*)?(static\s+)?(CB_DEV_NOINLINE|CB_DEV|CB_MEM|__global__|__device__|inline)\b.*\(")
name_re = re.compile(r"([A-Za-z_][A-Za-z0-9_]*)\s*\(")
files = {}
cur = None; hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1]; hdr = None; continue
    if len(r) >= 2 and r[0] == "Function Name": continue
    if r and r[0] == "Line No": hdr = r; continue
    if cur is None or hdr is None or not r or not r[0].isdigit(): continue
    files.setdefault(cur, []).append(r)
    files.setdefault((cur, "hdr"), hdr)
fn_inst = collections.Counter(); fn_samp = collections.Counter(); fn_thr = collections.Counter()
lines = []
for f, rs in files.items():
    if isinstance(f, tuple): continue
    h = files[(f, "hdr")]
    ii = h.index("Instructions Executed"); si = h.index("# Samples"); ti = h.index("Thread Instructions Executed")
    fn = "?"
    for r in rs:
        src = r[1]
        if hdr_re.match(src) and not src.strip().startswith("//"):
            pre = src.split("(")[0]
            m = name_re.findall(src.split("{")[0])
            cand = [x for x in m if x not in ("__launch_bounds__", "if", "for", "while", "defined")]
            if cand: fn = cand[0]
        try:
            n = int(r[ii].replace(",", "") or 0); s = int(r[si].replace(",", "") or 0); t = int(r[ti].replace(",", "") or 0)
        except ValueError:
            continue
        if n or s:
            key = "%s:%s" % (f.split("/")[-1], fn)
            fn_inst[key] += n; fn_samp[key] += s; fn_thr[key] += t
            lines.append((n, s, t, f.split("/")[-1], r[0], src.strip()[:100]))
ti = sum(fn_inst.values()) or 1; ts = sum(fn_samp.values()) or 1
print("total warp instructions %d, samples %d" % (ti, ts))
print("| function | inst %% | samples %% | lanes |\n|---|---|---|---|")
for k, n in fn_inst.most_common(45):
    print("| %s | %.1f | %.1f | %.1f |" % (k, 100.0 * n / ti, 100.0 * fn_samp[k] / ts, fn_thr[k] / max(1, n)))
print()
for n, s, t, f, ln, src in sorted(lines, key=lambda x: -x[1])[:topn]:
    print("%5.1f%% inst %5.1f%% samp %4.1f lanes  %s:%s  %s" % (100.0 * n / ti, 100.0 * s / ts, t / max(1, n), f, ln, src))
