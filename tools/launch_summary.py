"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time and
share of all GPU time in the run (cold-cache, serialised times: compare SHARES with bench.py, not absolutes)."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    rows.append((r["Kernel Name"].split("(")[0][:90], ns))
tot = sum(ns for _, ns in rows)
agg = defaultdict(lambda: [0, 0.0])
for k, ns in rows:
    agg[k][0] += 1
    agg[k][1] += ns
print(f"{len(rows)} launches, {tot/1e6:.1f} ms total GPU time")
print(f"{'share':>7} {'total ms':>10} {'launches':>8} {'avg us':>10}  kernel")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{100*ns/tot:6.2f}% {ns/1e6:10.2f} {n:8d} {ns/n/1e3:10.1f}  {k}")
