"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and mean duration, share."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]), v))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = defaultdict(lambda: [0, 0.0])
for k, v in rows:
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in rows)
print(f"{len(rows)} launches, {tot:.1f} us total device time (serialised, cold cache: compare SHARES)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={n:4d}  mean={t / n:8.2f} us  {k[:90]}")
