"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total us, share."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, agg, seq = None, collections.OrderedDict(), []
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = d["Kernel Name"][:64]
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
    seq.append((k, v))
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:64s} {a[0]:4d} {a[1]:9.1f} us {100 * a[1] / tot:5.1f}%  ({a[1] / a[0]:.1f} us each)")
print(f"total {tot:.1f} us over {len(seq)} launches")
if len(sys.argv) > 2:
    for k, v in seq:
        print(f"  {v:8.1f}  {k}")
