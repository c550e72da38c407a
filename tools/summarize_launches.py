"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time and share per kernel family.
    python tools/summarize_launches.py gpurun_out/launches.csv "header comment" > profiles/rNN_launches_summary.txt"""
import collections
import csv
import re
import sys

path = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 else ""
lines = [l for l in open(path) if l.startswith('"')]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    m = re.search(r"a3d::(\w+)(<[^>]*>)?", name)
    if not m:
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v / 1e6 if unit.startswith("n") else v / 1e3 if unit.startswith("u") else v
    key = m.group(1) + (m.group(2) or "")
    agg[key][0] += 1
    agg[key][1] += ms
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
print(f"# {note}")
print(f"# a3d kernels: {n} launches, {tot:.1f} ms serialised (per-launch times under ncu are cold-cache; the SHARES are what to compare)")
for k, (c, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{ms:9.3f} ms {100 * ms / tot:5.1f}% n={c:4d} {k}")
