"""DRAM bytes per frame of a kernel from an `ncu --set full` report (dram__bytes_read.sum + dram__bytes_write.sum of the captured
launch / frames of that launch) -> the JSON files bench.py reads for roofline.traffic.

usage: python tools/traffic_from_ncu.py out.json "source text" kernel=report.ncu-rep:frames ...
"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def dram_bytes(rep, kern):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    row = [r for r in rows[2:] if kern in dict(zip(hdr, r))["Kernel Name"]][-1]
    d, u = dict(zip(hdr, row)), dict(zip(hdr, units))
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(d[k].replace(",", "")) * UNIT[u[k]]
    return tot


def main():
    out, source = sys.argv[1], sys.argv[2]
    res = {}
    for spec in sys.argv[3:]:
        kern, rest = spec.split("=")
        rep, frames = rest.rsplit(":", 1)
        res[kern] = dram_bytes(rep, kern) / float(frames)
    json.dump({"source": source, "dram_bytes_per_frame": res}, open(out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
