"""Summarise an ncu --csv launch list (gpu__time_duration.sum) by kernel name: count, total, share."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v_us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    tot[name][0] += 1
    tot[name][1] += v_us
total = sum(v[1] for v in tot.values())
print(f"{'kernel':70s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:5d} {t:10.1f} {t / n:9.1f} {100 * t / total:5.1f}%")
print(f"{'TOTAL':70s} {sum(v[0] for v in tot.values()):5d} {total:10.1f}")
