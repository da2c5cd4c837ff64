"""Kernel share table of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r[4], [0, 0.0])
    a[0] += 1
    a[1] += float(r[-1])
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-52s n=%4d avg=%8.1f us share=%5.1f%%" % (k[:52], a[0], a[1] / a[0] / 1e3, 100 * a[1] / tot))
