"""Per-device-function breakdown of one kernel in an .ncu-rep (ncu --set full --import-source on).

The kernels are built "small-code" (medium helpers are real calls), so the kernel's .text holds the callee functions as
local FUNC symbols.  This tool reads those symbols from the cubin of the SAME build and attributes every SASS row of the
report's source page to the function it lies in.

usage: python tools/ncu_by_function.py report.ncu-rep file.cubin|lib.so kernel_substring [--md]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def cubin_symbols(path, kern):
    """[(offset, size, short name)] of the FUNC symbols inside the kernel's .text section."""
    if path.endswith(".so"):
        d = tempfile.mkdtemp()
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(path)], cwd=d, capture_output=True)
        cubins = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")]
    else:
        cubins = [path]
    out = []
    for cb in cubins:
        txt = subprocess.run(["cuobjdump", "-elf", cb], capture_output=True, text=True).stdout
        if kern not in txt:
            continue
        ksize = 0
        for ln in txt.splitlines():
            f = ln.split()
            if len(f) >= 7 and f[0].startswith("0x") and kern in f[-1]:
                if f[3] in ("0x12", "0x2") and not f[-1].startswith(".") and not f[-1].startswith("$"):
                    ksize = int(f[2], 16)   # kernels in an anonymous namespace are local FUNC symbols
                elif f[3] in ("0x2", "0x22") and f[-1].startswith("$") and "$" in f[-1][1:]:
                    name = f[-1].split("$")[-1]
                    out.append((int(f[1], 16), int(f[2], 16), name))
        if ksize:
            out.sort()
            first = out[0][0] if out else ksize
            out.insert(0, (0, first, "<kernel body + inlined>"))
            return out, ksize
    raise SystemExit("kernel %s not found in %s" % (kern, path))


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = []
    for n in r:
        n = re.sub(r"\[with.*", "", n)
        n = re.sub(r"\(.*", "", n)
        n = re.sub(r"<.*", "", n)
        n = n.replace("(anonymous namespace)::", "")
        n = re.sub(r"^.*?(\w+::)?(\w+)$", lambda m: (m.group(1) or "") + m.group(2), n.split(" ")[-1])
        short.append(n)
    return short


def main():
    rep, cub, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    syms, ksize = cubin_symbols(cub, kern)
    names = demangle([s[2] for s in syms])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # several kernels may be in the report: keep the LAST section whose name matches
    sections, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = [r[1], None, []]
            sections.append(cur)
        elif r and r[0] == "Address" and cur is not None:
            cur[1] = r
        elif r and r[0].startswith("0x") and cur is not None:
            cur[2].append(r)
    short = kern.split("ILi")[0]   # the report names kernels demangled, the cubin mangled
    sec = [s for s in sections if kern in s[0] or short in s[0]][-1]
    hdr, body = sec[1], sec[2]
    col = {h: i for i, h in enumerate(hdr)}
    base = int(body[0][0], 16)
    if len(body) * 16 != ksize:
        sys.stderr.write("warning: report has %d instructions, cubin kernel has %d: build mismatch?\n" % (len(body), ksize // 16))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.defaultdict(lambda: collections.Counter())
    starts = [s[0] for s in syms]
    import bisect
    for r in body:
        off = int(r[0], 16) - base
        k = bisect.bisect_right(starts, off) - 1
        a = agg[k]
        a["inst"] += int(r[col["Instructions Executed"]])
        a["thr"] += int(r[col["Thread Instructions Executed"]])
        a["samp"] += int(r[col["# Samples"]])
        a["n"] += 1
        for h in stall_cols:
            a[h] += int(r[col[h]] or 0)
        if r[col["Address Space"]] == "Local":
            a["local"] += int(r[col["Instructions Executed"]])
    ti = sum(a["inst"] for a in agg.values()) or 1
    ts = sum(a["samp"] for a in agg.values()) or 1
    print("| function | SASS inst | warp inst | share | samples | lanes/inst | local ld/st | top stalls |")
    print("|---|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"]):
        st = sorted(((a[h], h[6:]) for h in stall_cols), reverse=True)[:3]
        sts = ", ".join("%s %.0f%%" % (h, 100.0 * v / max(1, a["samp"])) for v, h in st if v)
        print("| %s | %d | %d | %.1f %% | %.1f %% | %.1f | %.1f %% | %s |" % (
            names[k], a["n"], a["inst"], 100.0 * a["inst"] / ti, 100.0 * a["samp"] / ts,
            a["thr"] / max(1, a["inst"]), 100.0 * a["local"] / max(1, a["inst"]), sts))
    print("\ntotal warp instructions %d, samples %d" % (ti, ts))


if __name__ == "__main__":
    main()
