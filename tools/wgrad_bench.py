"""Micro-benchmark of mq_conv_wgrad on the refiner's layer shapes at the training step's size (BASELINE configs[4]:
16 utterances x 256 frames per GPU).  Usage: python tools/wgrad_bench.py [layer-substring] [utterances]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mqgan_b200 import ops

only = sys.argv[1] if len(sys.argv) > 1 else ""
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
T, F = 256, 144
LAYERS = [  # name, Cin, Cout, time stride
    ("pre.conv2", 64, 64, 1), ("down0.conv1", 64, 128, 2), ("down0.conv2", 128, 128, 2), ("down1.conv1", 128, 256, 4),
    ("down1.conv2", 256, 256, 4), ("down2.conv1", 256, 512, 8), ("down2.conv2", 512, 512, 8), ("mid.conv1", 512, 512, 8),
    ("up0.conv1", 768, 256, 4), ("up0.conv2", 256, 256, 4), ("up1.conv1", 384, 128, 2), ("up1.conv2", 128, 128, 2),
    ("up2.conv1", 192, 64, 1), ("up2.conv2", 64, 64, 1),
]
dh, dw = ops.conv_taps("conv2d3", (1, 1, 3, 3))
for name, cin, cout, s in LAYERS:
    if only and only not in name:
        continue
    H = T // s
    x = torch.randn(B, H, F, cin, device="cuda").to(torch.bfloat16)
    dy = torch.randn(B, H, F, cout, device="cuda").to(torch.bfloat16)
    for _ in range(3):
        ops.conv_wgrad(dy, x, B, H, F, cout, cin, dh, dw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        ops.conv_wgrad(dy, x, B, H, F, cout, cin, dh, dw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * B * H * F * cout * cin * 9
    print(f"{name:14s} {cin:4d}->{cout:4d} 1/{s}  {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s (incl. the partial-sum reduction)", flush=True)
