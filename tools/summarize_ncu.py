"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, mean, share.
Usage: python tools/summarize_ncu.py gpurun_out/launches.csv [out.md]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = None
for r in rd:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    rows.append(r)
ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size") if "Grid Size" in hdr else None
agg = defaultdict(lambda: [0, 0.0])
order = []
for r in rows:
    name = re.sub(r"\(.*\)$", "", r[ki]).replace("void ", "").replace("b2d::", "")
    v = float(r[vi].replace(",", ""))
    unit = r[ui]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    agg[name][0] += 1
    agg[name][1] += us
    order.append((name, us))
tot = sum(v[1] for v in agg.values())
out = [f"launches: {len(rows)}   total kernel time: {tot/1e3:.3f} ms (serialised, cold-cache: compare shares)\n",
       "| kernel | launches | total ms | mean us | share |", "|---|---:|---:|---:|---:|"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {n} | {t/1e3:.3f} | {t/n:.1f} | {100*t/tot:.1f}% |")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
