"""Where the warps of conv_pair_kernel spend their cycles on a narrow refiner layer (probe build: `make -C mqgan_b200/csrc probes`,
run with MQ_LIB=mqgan_b200/libmqgan_b200_probes.so MQ_SLIM16=0).  clock64() deltas summed over lane 0 of every epilogue
warp / of the leader's MMA warp, divided by the number of (warp, tile) passes.
Usage: python tools/conv_cycles.py [shape filter]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mqgan_b200 import _lib, ops

lib = _lib.lib()
lib.mq_conv_probe_cycles.restype = C.c_int
lib.mq_conv_probe_cycles.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
SHAPES = [("pre.conv2 64->64 + pool", 1024, 144, 64, 64, True), ("up2.conv2 64->64", 1024, 144, 64, 64, False),
          ("down0.conv1 64->128 + pool", 512, 144, 64, 128, True), ("down0.conv2 128->128", 512, 144, 128, 128, False),
          ("mid.conv1 512->512", 128, 144, 512, 512, False),
          ("up2.conv1 128+64->64 (fused up-concat)", 1024, 144, (128, 64), 64, False),
          ("up1.conv1 256+128->128 (fused up-concat)", 512, 144, (256, 128), 128, False)]
FILT = sys.argv[1] if len(sys.argv) > 1 else ""
B, dev = 32, "cuda"
buf = (C.c_ulonglong * 12)()
for name, H, W, Cin, Cout, pool in SHAPES:
    if FILT and FILT not in name:
        continue
    y = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=dev)
    mask = torch.zeros(B, H, dtype=torch.uint8, device=dev)
    if isinstance(Cin, tuple):
        Cx, Cs = Cin
        x = torch.randn(B, H // 2, W, Cx, device=dev).to(torch.bfloat16)
        skip = torch.randn(B, H, W, Cs, device=dev).to(torch.bfloat16)
        w = torch.randn(Cout, Cx + Cs, 3, 3) / (9 * (Cx + Cs)) ** 0.5
        pc = ops.pack_upconv(w, torch.zeros(Cout), Cx, Cs).to(dev)

        def run():
            ops.conv_gemm(x, pc, B, H // 2, W, x2=skip, act=True, row_mask=mask, mask_post=True, out_bf16=y, pair=True)
    else:
        x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
        w = torch.randn(Cout, Cin, 3, 3) / (9 * Cin) ** 0.5
        pc = ops.pack_conv(w, torch.zeros(Cout), "conv2d3", False).to(dev)
        yp = torch.empty(B, H // 2, W, Cout, dtype=torch.bfloat16, device=dev) if pool else None

        def run():
            ops.conv_gemm(x, pc, B, H, W, act=True, row_mask=mask, mask_post=True, out_bf16=y, pair=True, out_pool=yp)

    for _ in range(3):
        run()
    lib.mq_conv_probe_cycles(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 5
    for _ in range(n):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    lib.mq_conv_probe_cycles(buf, 1)
    v = [float(buf[i]) for i in range(12)]
    passes = max(v[4], 1.0)                      # (epilogue warp, tile) passes over n launches
    tiles_mma = passes / 16.0                    # 8 epilogue warps x 2 CTAs per pair tile
    print(f"{name:28s} {ms:.3f} ms | per (epilogue warp, tile): wait accumulator {v[0] / passes:8.0f}  tcgen05.ld {v[1] / passes:7.0f}  "
          f"body {v[2] / passes:8.0f}  total {v[3] / passes:8.0f} clk | MMA warp per tile: wait free accumulator {v[5] / tiles_mma:8.0f}  "
          f"wait activations {v[6] / tiles_mma:8.0f}  wait weights {v[8] / tiles_mma:8.0f}  issue {v[9] / tiles_mma:8.0f}  commits {v[10] / tiles_mma:8.0f}  total {v[7] / tiles_mma:8.0f} clk", flush=True)
