"""Which cuDNN path do the discriminators' strided convolutions take?  Times forward + backward (data and weight
gradients) of every conv shape of the hifispeech discriminators under several formulations (development aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from mqgan_b200 import spec as S

B, T = 16, 256


def shapes(dc, mel):
    out, h, w = [], mel, T
    for i, (co, ci, kh, kw) in enumerate(dc.conv_shapes()):
        sh, sw = dc.layer_stride(i)
        out.append((ci, co, kh, kw, sh, sw, h, w))
        h, w = -(-h // sh), -(-w // sw)
    return out


def run(variant, ci, co, kh, kw, sh, sw, h, w):
    x = torch.randn(B, ci, h, w, device="cuda", requires_grad=True)
    wt = (torch.randn(co, ci, kh, kw, device="cuda") * 0.02).requires_grad_(True)
    pad = ((kh - 1) // 2, (kw - 1) // 2)
    cl = "cl" in variant
    bf = "bf16" in variant
    sub = "sub" in variant

    def f():
        xi, wi = x, wt
        if cl:
            xi, wi = xi.contiguous(memory_format=torch.channels_last), wi.contiguous(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf):
            if sub and (sh > 1 or sw > 1):
                y = F.conv2d(xi, wi, None, stride=1, padding=pad)[:, :, ::sh, ::sw]
            else:
                y = F.conv2d(xi, wi, None, stride=(sh, sw), padding=pad)
        y.float().sum().backward()
    for _ in range(2):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3


variants = ["bf16", "bf16_cl", "bf16_sub", "bf16_cl_sub", "fp32_tf32", "fp32_tf32_cl"]
for name, dc, mel in (("patch", S.HIFISPEECH_PATCH_D, 128), ("multibin", S.HIFISPEECH_MULTIBIN_D.bin_config, 16)):
    for shp in shapes(dc, mel):
        ms = [run(v, *shp) for v in variants]
        print(name, shp, " ".join(f"{v}={m:.3f}" for v, m in zip(variants, ms)), flush=True)
