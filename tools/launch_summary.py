"""Totals per kernel of an ncu launch list (--csv --metrics gpu__time_duration.sum).
python tools/launch_summary.py launches.csv ["header line"]"""
import csv
import re
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
tot, cnt = defaultdict(float), defaultdict(int)
for r in csv.DictReader(rows):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("void ", "")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[r["Metric Unit"]]
    tot[name] += float(r["Metric Value"].replace(",", "")) * scale
    cnt[name] += 1
total = sum(tot.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
print(f"total {total:.1f} ms over {sum(cnt.values())} launches")
for name, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"  {ms:8.2f} ms {100 * ms / total:5.1f} %  n={cnt[name]:5d}  {name}")
