#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total and mean duration, share.
    python tools/launch_summary.py gpurun_out/launches.csv"""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void |xb::|\(anonymous namespace\)::", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit.startswith("ns") else (v if unit.startswith("us") else v * 1e3 if unit.startswith("ms") else v)
    tot[name][0] += 1
    tot[name][1] += us
all_us = sum(v[1] for v in tot.values())
print(f"{'kernel':70s} {'launches':>9s} {'total us':>12s} {'mean us':>10s} {'share':>7s}")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:9d} {us:12.1f} {us / n:10.2f} {100 * us / all_us:6.1f}%")
print(f"{'TOTAL':70s} {sum(v[0] for v in tot.values()):9d} {all_us:12.1f}")
