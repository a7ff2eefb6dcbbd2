"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.
Usage: python tools/summarize_launches.py launches.csv [title] > summary.md"""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows:
    if r is hdr or len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(r[iu], 1e-6)
    name = r[ik].split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v * scale
tot = sum(a[1] for a in agg.values())
n = sum(a[0] for a in agg.values())
print(f"# {title}\n")
print("Per-launch times are cold-cache and serialised under ncu: compare SHARES with bench.py's CUDA-event table.\n")
print("| kernel | launches | ms | share |\n|---|---|---|---|")
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {c} | {ms:.3f} | {100 * ms / tot:.1f}% |")
print(f"\ntotal {tot:.3f} ms over {n} launches")
