"""Micro-benchmark of mq_vq_nearest (BASELINE configs[3]): N = 2^20 latents against codebooks of
1024 / 8192 codes; D = 4 / 5 (FSQ implicit codebooks, the reference's only quantiser) and D = 64
(random normal codebook).  Prints time, 2NKD TFLOP/s, algorithmic HBM bytes and GB/s, and checks a
sample of rows against a float64 argmin."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mqgan_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = "cuda"
CASES = [("fsq[8,8,4,4]", [8, 8, 4, 4], None), ("fsq[8,8,8,4,4]", [8, 8, 8, 4, 4], None),
         ("random K=1024 D=64", None, (1024, 64)), ("random K=8192 D=64", None, (8192, 64))]
out = []
for name, levels, shape in CASES:
    g = torch.Generator().manual_seed(0)
    if levels is not None:
        K, D = int(np.prod(levels)), len(levels)
        lv = torch.tensor(levels)                                   # FSQ implicit codebook (quantizer.py:101-104, 183-187)
        basis = torch.cumprod(torch.tensor([1] + list(levels[:-1])), dim=0)
        cb = ((((torch.arange(K)[:, None] // basis) % lv) - lv // 2) / (lv // 2)).float()
        z = (torch.randn(N, D, generator=g) * 0.6).clamp(-1.05, 1.05)
    else:
        K, D = shape
        cb = torch.randn(K, D, generator=g)
        z = torch.randn(N, D, generator=g)
    zd = z.to(dev)
    for prec in ("f16x2", "bf16"):
        pc = ops.pack_codebook(cb, prec).to(dev)
        for _ in range(3):
            idx, codes = ops.vq_nearest(zd, pc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        reps = 10
        for _ in range(reps):
            idx, codes = ops.vq_nearest(zd, pc)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        sample = torch.arange(0, N, max(1, N // 4096))[:4096]
        d = torch.cdist(z[sample].double(), cb.double())
        ref = torch.argmin(d, dim=1)
        agree = float((idx.cpu()[sample] == ref).float().mean())
        bytes_alg = N * D * 4 + N * (8 + D * 4)
        rec = {"case": name, "precision": prec, "n": N, "k": K, "d": D, "ms": ms,
               "tflops_2nkd": 2.0 * N * K * D / ms / 1e9, "alg_bytes": bytes_alg, "gbs": bytes_alg / ms / 1e6,
               "index_agreement_vs_fp64_sample": agree}
        out.append(rec)
        print(json.dumps(rec), flush=True)
