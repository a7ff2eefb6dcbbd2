"""Micro-benchmark of mq_conv_gemm on the refiner layer shapes (bottleneck probes).
Usage: python tools/conv_bench.py [B] ; env MQ_CONV_DEBUG / MQ_MSUB / MQ_CONV_STAGES select variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mqgan_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
SHAPES = [  # name, H, W, Cin, Cout
    ("pre.conv2 64->64 L0", 1024, 144, 64, 64),
    ("down0.conv1 64->128 L1", 512, 144, 64, 128),
    ("down0.conv2 128->128 L1", 512, 144, 128, 128),
    ("down1.conv2 256->256 L2", 256, 144, 256, 256),
    ("up1.conv1 384->128 L1", 512, 144, 384, 128),
    ("up2.conv1 192->64 L0", 1024, 144, 192, 64),
    ("mid.conv1 512->512 L3", 128, 144, 512, 512),
]
dev = "cuda"
for name, H, W, Cin, Cout in SHAPES:
    x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    w = torch.randn(Cout, Cin, 3, 3) / (9 * Cin) ** 0.5
    pc = ops.pack_conv(w, torch.zeros(Cout), "conv2d3", False).to(dev)
    y = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=dev)
    mask = torch.zeros(B, H, dtype=torch.uint8, device=dev)
    def run():
        ops.conv_gemm(x, pc, B, H, W, act=True, row_mask=mask, mask_post=True, out_bf16=y)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    n = 10
    for _ in range(n):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 2.0 * B * H * W * Cout * Cin * 9
    print(f"{name:28s} {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s  msub={ops.choose_msub(pc.bn, B, H, W, *ops.choose_tile(H, W))}", flush=True)
