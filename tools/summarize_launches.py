"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time and share.
Usage: python tools/summarize_launches.py launches.csv > summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
reader = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in reader:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    unit = r.get("Metric Unit", "ns")
    v = float(r["Metric Value"].replace(",", ""))
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    tot[name][0] += 1
    tot[name][1] += us
total = sum(v[1] for v in tot.values())
n = sum(v[0] for v in tot.values())
for name, (cnt, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:72]:72s} n={cnt:5d} us={us:11.1f} share={us / total:6.3f}")
print(f"total us {total:.1f} over {n} launches (ncu: cold-cache, serialised; compare shares)")
