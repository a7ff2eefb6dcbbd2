"""Turn the ncu launch list of one step (tools/ncu_step.py under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,
gpu__time_duration.sum --csv`) into profiles/ncu_traffic_r02.json: DRAM bytes per launch per kernel family, which bench.py
reports as roofline.traffic (with its source) instead of a literal.

    python tools/ncu_traffic.py gpurun_out/traffic.csv hifispeech [algorithmic_bytes_per_launch] > summary.md
"""
import csv
import json
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, model = sys.argv[1], sys.argv[2]
alg = float(sys.argv[3]) if len(sys.argv) > 3 else None
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
iid, ik, im, iv, iu = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}
per = OrderedDict()
for r in rows:
    if r is hdr or len(r) <= iv:
        continue
    k = per.setdefault(r[iid], {"name": r[ik].split("(")[0].replace("void ", "")})
    k[r[im]] = float(r[iv].replace(",", "")) * UNIT.get(r[iu], 1.0)
agg = OrderedDict()
for k in per.values():
    a = agg.setdefault(k["name"], [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += k.get("dram__bytes_read.sum", 0.0)
    a[2] += k.get("dram__bytes_write.sum", 0.0)
    a[3] += k.get("gpu__time_duration.sum", 0.0)
conv = [v for n, v in agg.items() if "conv_pair" in n or "conv_gemm_kernel" in n or "conv_halo" in n]
out_path = os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")
data = json.load(open(out_path)) if os.path.exists(out_path) else {}
n = sum(v[0] for v in conv)
entry = {"launches": n, "dram_bytes_total": sum(v[1] + v[2] for v in conv),
         "dram_bytes_per_launch": sum(v[1] + v[2] for v in conv) / max(n, 1),
         "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum over all {n} conv launches of one step ({os.path.basename(path)}, "
                   "tools/ncu_step.py); per-launch average like roofline.achieved"}
if alg is not None:
    entry["algorithmic_bytes_per_launch"] = alg
data.setdefault(model, {})["mq_conv_gemm"] = entry
data[model]["kernels"] = {nme: {"launches": v[0], "dram_read_bytes": v[1], "dram_write_bytes": v[2], "ms_under_ncu": v[3]} for nme, v in agg.items()}
json.dump(data, open(out_path, "w"), indent=1)
tot = sum(v[3] for v in agg.values())
print(f"# DRAM traffic and time per kernel, one step ({model}), from {os.path.basename(path)}\n")
print("| kernel | launches | DRAM read MB | DRAM write MB | ms (ncu, cold, serialised) | share |\n|---|---|---|---|---|---|")
for nme, v in sorted(agg.items(), key=lambda kv: -kv[1][3]):
    print(f"| {nme} | {v[0]} | {v[1] / 1e6:.1f} | {v[2] / 1e6:.1f} | {v[3]:.3f} | {100 * v[3] / tot:.1f}% |")
print(f"\ntotal {tot:.3f} ms; conv family {n} launches, {entry['dram_bytes_per_launch'] / 1e6:.1f} MB of DRAM traffic per launch")
