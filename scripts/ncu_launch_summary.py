"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares.
usage: python scripts/ncu_launch_summary.py launches.csv [fraction_of_tail_to_keep]"""
import collections
import csv
import sys

path = sys.argv[1]
tail = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
    seq.append((row["Kernel Name"][:90], v))
seq = seq[int(len(seq) * (1 - tail)):]
agg = collections.OrderedDict()
for name, v in seq:
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v for _, v in seq)
print(f"{len(seq)} launches, {tot:.1f} us of kernel time (cold-cache, serialised: compare shares)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={n:4d}  avg={t / n:9.1f}  {k}")
