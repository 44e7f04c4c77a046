#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

lines = open(sys.argv[1]).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(lines[start:]))
agg = collections.OrderedDict()
for r in rows:
    k = r['Kernel Name'].split('(')[0].replace('<unnamed>::', '')
    a = agg.setdefault((k, r['Grid Size'], r['Block Size']), [0, 0.0])
    a[0] += 1
    a[1] += float(r['Metric Value']) / 1e3
tot = sum(a[1] for a in agg.values())
print("| kernel | grid | block | launches | total us | mean us | share |\n|---|---|---|---|---|---|---|")
for (k, g, b), a in agg.items():
    print("| %s | %s | %s | %d | %.1f | %.1f | %.1f %% |" % (k, g, b, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
print("total %.1f us" % tot)
