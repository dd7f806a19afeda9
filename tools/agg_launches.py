"""aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = row["Kernel Name"][:int(sys.argv[2]) if len(sys.argv) > 2 else 60]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    a = agg.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
print(f"total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches")
for k, (n, t, m) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
    print(f"{k:60s} n={n:4d} total={t:9.3f} ms ({100*t/tot:5.1f}%) avg={t/n:8.3f} max={m:8.3f}")
