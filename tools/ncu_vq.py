"""One mq_vq_nearest launch between cudaProfilerStart / Stop (ncu --profile-from-start off ...).
Usage: python tools/ncu_vq.py [K] [D] [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mqgan_b200 import ops

K = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
D = int(sys.argv[2]) if len(sys.argv) > 2 else 5
prec = sys.argv[3] if len(sys.argv) > 3 else "f16x2"
g = torch.Generator().manual_seed(0)
cb = torch.randn(K, D, generator=g)
z = torch.randn(1 << 20, D, generator=g).cuda()
pc = ops.pack_codebook(cb, prec).to("cuda")
for _ in range(2):
    ops.vq_nearest(z, pc)
torch.cuda.synchronize()
torch.cuda.profiler.start()
idx, codes = ops.vq_nearest(z, pc)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("vq", K, D, prec, "fold", pc.fold, int(idx.sum()))
