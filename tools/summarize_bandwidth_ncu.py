"""ncu --csv launch list with gpu__time_duration.sum + dram__bytes_{read,write}.sum -> per-kernel table: launches,
median duration, DRAM bytes moved per launch, achieved DRAM GB/s and its share of the measured copy peak.
    python tools/summarize_bandwidth_ncu.py gpurun_out/x.csv [name-regex] > profiles/x.txt"""
import csv
import json
import os
import re
import statistics
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6541.1
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
flt = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
per = defaultdict(dict)     # launch id -> metric -> value
names = {}
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", ""))
    u = r.get("Metric Unit", "")
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        v = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)       # -> us
    elif m.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    per[r["ID"]][m] = v
    names[r["ID"]] = re.sub(r"\(.*", "", r["Kernel Name"])
agg = defaultdict(list)
for i, d in per.items():
    if flt and not flt.search(names[i]):
        continue
    if "gpu__time_duration.sum" in d:
        agg[names[i]].append((d["gpu__time_duration.sum"], d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)))
print(f"# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; peak = measured copy {peak:.0f} GB/s")
print(f"{'kernel':56s} {'n':>4s} {'med us':>9s} {'read MB':>9s} {'write MB':>9s} {'GB/s':>8s} {'% peak':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(x[0] for x in kv[1])):
    us = statistics.median(x[0] for x in v)
    rd = statistics.median(x[1] for x in v)
    wr = statistics.median(x[2] for x in v)
    gbs = (rd + wr) / (us * 1e-6) / 1e9
    print(f"{k[:56]:56s} {len(v):4d} {us:9.1f} {rd / 1e6:9.1f} {wr / 1e6:9.1f} {gbs:8.0f} {100 * gbs / peak:6.1f}%")
