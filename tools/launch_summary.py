"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/launch_summary.py profiles/r02_bench_launches.csv
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = d["Kernel Name"].split("(")[0][:70]
    val = float(d["Metric Value"].replace(",", ""))
    val = val / 1000.0 if d["Metric Unit"] == "ns" else (val * 1000.0 if d["Metric Unit"] == "ms" else val)
    a = agg.setdefault(name, [0, 0.0, 1e30, 0.0])
    a[0] += 1; a[1] += val; a[2] = min(a[2], val); a[3] = max(a[3], val)
tot = sum(a[1] for a in agg.values())
print("%-72s %6s %12s %9s %9s %6s" % ("kernel", "n", "total us", "min us", "max us", "share"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %6d %12.1f %9.1f %9.1f %5.1f%%" % (k, a[0], a[1], a[2], a[3], 100 * a[1] / tot))
