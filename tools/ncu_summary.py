"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if u == "ns":
        v /= 1e3
    elif u == "ms":
        v *= 1e3
    elif u == "s":
        v *= 1e6
    a = agg[r[ki][:48]]
    a[0] += 1
    a[1] += v
    a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':48s} {'launches':>8s} {'total_us':>10s} {'share':>7s} {'avg_us':>9s} {'max_us':>9s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print(f"{k:48s} {a[0]:8d} {a[1]:10.1f} {a[1] / tot:7.1%} {a[1] / a[0]:9.2f} {a[2]:9.1f}")
