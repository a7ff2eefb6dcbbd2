"""Micro-benchmark of mq_conv_gemm on the encoder / decoder 1-D layer shapes: tap-shifted single-CTA loop vs
the row-halo CTA-pair loop.  Usage: python tools/conv1d_bench.py [rows]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mqgan_b200 import ops

ROWS = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
T = 1024
B = ROWS // T
SHAPES = [  # name, kind, Cin, Cout, k, precision
    ("enc0.conv1 512->512 k3 f16x2", "same1d", 512, 512, 3, "f16x2"),
    ("enc2.conv1 512->768 k5 f16x2", "same1d", 512, 768, 5, "f16x2"),
    ("enc2.conv2 768->768 k5 f16x2", "same1d", 768, 768, 5, "f16x2"),
    ("enc2.res 512->768 k1 f16x2", "linear", 512, 768, 1, "f16x2"),
    ("dec0.conv1 768->512 k7 bf16", "causal1d", 768, 512, 7, "bf16"),
    ("dec1.conv1 512->512 k5 bf16", "causal1d", 512, 512, 5, "bf16"),
    ("dec2.conv1 512->512 k3 bf16", "causal1d", 512, 512, 3, "bf16"),
]
dev = "cuda"
for name, kind, cin, cout, k, prec in SHAPES:
    w = torch.randn(*((cout, cin) if kind == "linear" else (cout, cin, k))) / (cin * k) ** 0.5
    pc = ops.pack_conv(w, torch.zeros(cout), kind, split=prec).to(dev)
    nt = ops.SPLIT_TERMS[prec]
    x = torch.randn(B * T, nt * cin, device=dev).to(torch.float16 if prec == "f16x2" else torch.bfloat16)
    if prec == "bf16":
        y = torch.empty(B * T, cout, dtype=torch.bfloat16, device=dev)
        kw = {"out_bf16": y}
    else:
        y = torch.empty(B * T, cout, dtype=torch.float32, device=dev)
        kw = {"out_f32": y}
    for pair in (False, True):
        def run():
            ops.conv_gemm(x, pc, B, T, 1, act=True, pair=pair, **kw)
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * B * T * cout * cin * k * pc.nseg
        print(f"{name:32s} {'pair' if pair else 'tap ':5s} {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s issued", flush=True)
