"""Micro-benchmark of mq_conv_gemm on the refiner layer shapes, one line per (layer, main loop, msub).
Usage: python tools/conv_bench.py [B] [filter]
Main loops: halo = single-CTA halo tile, pair = CTA pair (cta_group::2), tap = tap-shifted (fused up-conv only).
Env MQ_CONV_DEBUG / MQ_CONV_STAGES select bottleneck probes of the single-CTA kernels; CONV_BENCH_SUSTAIN=S times each
variant inside S seconds of back-to-back launches (power-capped clocks, as inside the step) and adds the energy per launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mqgan_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
FILT = sys.argv[2] if len(sys.argv) > 2 else ""
# name, H (output rows), W, Cin (or (Cx, Cs) for the fused up-conv), Cout
SHAPES = [
    ("pre.conv2 64->64 L0", 1024, 144, 64, 64),
    ("down0.conv1 64->128 L1", 512, 144, 64, 128),
    ("down0.conv2 128->128 L1", 512, 144, 128, 128),
    ("down1.conv1 128->256 L2", 256, 144, 128, 256),
    ("down1.conv2 256->256 L2", 256, 144, 256, 256),
    ("down2.conv1 256->512 L3", 128, 144, 256, 512),
    ("mid.conv1 512->512 L3", 128, 144, 512, 512),
    ("up0.conv1 512+256->256 L2", 256, 144, (512, 256), 256),
    ("up1.conv1 256+128->128 L1", 512, 144, (256, 128), 128),
    ("up2.conv1 128+64->64 L0", 1024, 144, (128, 64), 64),
]
dev = "cuda"


SUSTAIN = float(os.environ.get("CONV_BENCH_SUSTAIN", "0"))     # seconds of back-to-back launches before (and while) timing
_nvml = None
if SUSTAIN > 0:
    try:
        import pynvml
        pynvml.nvmlInit()
        _nvml = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
    except Exception:  # noqa: BLE001 - energy column is optional
        _nvml = None
LAST = {}


def bench(fn, n=10):
    """ms per launch.  Default: 10 launches after 3 warm-ups (boost clocks).  CONV_BENCH_SUSTAIN=S: S seconds of
    back-to-back launches first, so the power governor settles where it sits inside the step; the time is taken
    over the last half of that run, with the board's energy counter around the same launches."""
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if SUSTAIN <= 0:
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    import time
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    per = max(e0.elapsed_time(e1) / 5, 1e-3)
    half = max(20, int(SUSTAIN * 500.0 / per))
    for _ in range(half):
        fn()
    torch.cuda.synchronize()
    j0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(_nvml) if _nvml is not None else 0
    t0 = time.perf_counter()
    e0.record()
    for _ in range(half):
        fn()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    j1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(_nvml) if _nvml is not None else 0
    LAST["mj_per_launch"] = (j1 - j0) / half if _nvml is not None else None
    LAST["watts"] = (j1 - j0) / 1e3 / wall if _nvml is not None else None
    return e0.elapsed_time(e1) / half


for name, H, W, Cin, Cout in SHAPES:
    if FILT and not any(f in name for f in FILT.split(",")):
        continue
    up = isinstance(Cin, tuple)
    mask = torch.zeros(B, H, dtype=torch.uint8, device=dev)
    y = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=dev)
    if up:
        Cx, Cs = Cin
        x = torch.randn(B, H // 2, W, Cx, device=dev).to(torch.bfloat16)
        skip = torch.randn(B, H, W, Cs, device=dev).to(torch.bfloat16)
        w = torch.randn(Cout, Cx + Cs, 3, 3) / (9 * (Cx + Cs)) ** 0.5
        pc = ops.pack_upconv(w, torch.zeros(Cout), Cx, Cs).to(dev)
        fl = 2.0 * B * H * W * Cout * (Cx + Cs) * 9
        variants = [("tap", None)] + [("pair", m) for m in (1, 2) if m * pc.bn <= 512]

        def run(mode, m):
            ops.conv_gemm(x, pc, B, H // 2, W, x2=skip, act=True, out_bf16=y, msub=m, pair=(mode == "pair"))
    else:
        x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
        w = torch.randn(Cout, Cin, 3, 3) / (9 * Cin) ** 0.5
        pc = ops.pack_conv(w, torch.zeros(Cout), "conv2d3", False).to(dev)
        fl = 2.0 * B * H * W * Cout * Cin * 9
        variants = [("halo", None)] + [("pair", m) for m in (1, 2, 4) if m * pc.bn <= 512]

        yp = (torch.empty(B, H // 2, W, Cout, dtype=torch.bfloat16, device=dev)
              if os.environ.get("CONV_BENCH_POOL") == "1" and Cout % 32 == 0 else None)      # fused AvgPool epilogue, as in the step

        def run(mode, m):
            ops.conv_gemm(x, pc, B, H, W, act=True, row_mask=mask, mask_post=True, out_bf16=y, msub=m,
                          pair=(mode == "pair"), halo=(mode == "halo"), out_pool=yp)
    for mode, m in variants:
        try:
            ms = bench(lambda: run(mode, m))
            extra = ""
            if SUSTAIN > 0 and LAST.get("mj_per_launch") is not None:
                extra = f"  {LAST['mj_per_launch']:8.1f} mJ/launch  {LAST['watts']:6.0f} W  sustained"
            print(f"{name:28s} {mode:5s} msub={str(m):5s} {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s{extra}", flush=True)
        except Exception as e:  # noqa: BLE001 - report and continue with the next variant
            print(f"{name:28s} {mode:5s} msub={str(m):5s} FAILED: {e}", flush=True)
