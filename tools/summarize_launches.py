"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share of ONE
step of bench.py (the launches between two consecutive patchify kernels).  Usage: summarize_launches.py in.csv"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    rows.append((r["Kernel Name"], ns))
starts = [i for i, (n, _) in enumerate(rows) if "patchify" in n]
if len(starts) < 3:
    sys.exit("need at least 3 steps in the capture")
a, b = starts[1], starts[2]          # the second step (first is cold)
step = rows[a:b]


def short(n):
    n = re.sub(r"\(.*", "", n)
    n = n.replace("apla::", "").replace("(anonymous namespace)::", "").replace("void ", "")
    return n[:70]


agg = OrderedDict()
for n, ns in step:
    k = short(n)
    c, t = agg.get(k, (0, 0.0))
    agg[k] = (c + 1, t + ns)
tot = sum(t for _, t in agg.values())
print(f"one step: {len(step)} launches, {tot / 1e6:.3f} ms of kernel time (serialised, cold-cache ncu replay)")
print(f"{'kernel':72s} {'n':>4s} {'total us':>10s} {'share':>7s}")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} {c:4d} {t / 1e3:10.1f} {100 * t / tot:6.1f}%")
