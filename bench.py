#!/usr/bin/env python
"""Benchmark of the PreEncoder re-encode pass (encode + FSQ + decode) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the UNMODIFIED reference on the host cores

One "step" = one re-encode pass over one batch of synthetic mels of the workload BASELINE.json's metric is quoted
on (configs[1]: hifispeech PreEncoder, 256 x 1024 frames, fp32-grade encoder so the VQ indices equal the fp32
reference; bf16 decoder / refiner).  Rank 0 prints ONE JSON line (contract in the task statement).  For N > 1 every
rank runs the same per-GPU workload on its own shard of utterances (weak scaling, no data-path collective; NCCL only
for the timing barrier and the max-over-ranks reduction).

The same line carries, outside the headline's timed region:
  parity     raw index agreement over ALL frames of the workload against a float64 restatement run on the GPU
             (oracle/gpu_checker.py, a checker), mel error on a 16-utterance slice, and agreement with the real
             reference's own fp32 indices on the CPU sample;
  secondary  the other BASELINE configs, each timed like the headline: the fp32-grade decoder mode of configs[1],
             configs[2] hifimusic 32 x 8192 bf16 with its index-agreement rate, configs[3] the VQ lookup microbench
             with a stated roofline per case, configs[4] the training step (tools/train_bench.py);
  cpu_baseline  the reference (baseline/_ref, kind "reference") or, without it, the oracle port (kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "mel frames/sec re-encoded (encode+VQ+decode)"
WORKLOADS = {
    # name: (config attr, B per GPU, T, encoder precision, decoder precision)
    "hifispeech_256x1024_fp32idx": ("HIFISPEECH", 256, 1024, "f16x2", "bf16"),
    "hifispeech_256x1024_fp32": ("HIFISPEECH", 256, 1024, "f16x2", "f16x2"),
    "hifimusic_32x8192_bf16": ("HIFIMUSIC", 32, 8192, "bf16", "bf16"),
    "hifispeech_16x512_fp32idx": ("HIFISPEECH", 16, 512, "f16x2", "bf16"),
    "tiny_8x256": ("TINY", 8, 256, "f16x2", "bf16"),
}
DEFAULT_WORKLOAD = "hifispeech_256x1024_fp32idx"
CPU_CONFIG0 = (16, 512)      # BASELINE configs[0]: the reference's own CPU-runnable case
PRECISION_TEXT = {
    "f16x2": "2-term fp16 operand splits (22-bit operands, three products per GEMM, fp32-grade)",
    "bf16x3": "3-term bf16 operand splits (24-bit operands, six products per GEMM, fp32-grade)",
    "bf16": "bf16 operands (single product)",
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], 0.0, set(), 0.0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = max(mx, float(parts[1]))
                power = max(power, float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "power_w_max": power or None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_model(cfg, precision, device, seed=0, decoder_precision="bf16"):
    """Random-init weights of the reference architecture + the SURVEY-D4 q_in_proj
    recalibration, computed with the CUDA encoder itself on a fixed calibration batch."""
    from mqgan_b200.preencoder import PreEncoder
    from mqgan_b200.synth import synth_state_dict, synth_mels, recalibrate_q_in_proj

    sd = synth_state_dict(cfg, seed=seed)
    extra = {} if decoder_precision == "bf16" else {"decoder_precision": decoder_precision}

    def make(sd_):
        m = PreEncoder(cfg.mel_channels, list(cfg.channels), list(cfg.kernel_sizes), fsq_levels=list(cfg.fsq_levels),
                       dropout=0.0, refiner_base_channels=cfg.refiner_base_channels, refiner_depth=cfg.refiner_depth,
                       refiner_hidden_proj_divisor=cfg.refiner_hidden_proj_divisor, encoder_precision=precision, **extra)
        m.load_state_dict(sd_, strict=True)
        return m.to(device).eval()

    m = make(sd)
    cal = synth_mels(8, 512, cfg.mel_channels, seed=100).to(device)
    _, z = m.engine().encode(cal, None, return_latents=True)
    recalibrate_q_in_proj(sd, z.cpu())
    return make(sd), sd


def calibrated_state_dict(cfg, seed=0):
    """The same calibrated state-dict without a GPU: latents of the calibration batch from the oracle port (used by
    the reference arm, which must not touch the CUDA path)."""
    from mqgan_b200.synth import synth_state_dict, synth_mels, recalibrate_q_in_proj
    from oracle import preencoder_oracle as O
    sd = synth_state_dict(cfg, seed=seed)
    cal = synth_mels(8, 512, cfg.mel_channels, seed=100)
    with torch.no_grad():
        z = O.encode_latents(sd, cfg, cal[:2, :128], None)
    recalibrate_q_in_proj(sd, z)
    return sd


def config_dict(workload, world):
    """The ``config`` object of the JSON line - identical in both arms (ours / --impl reference)."""
    cfg_name, B, T, precision, dec_precision = WORKLOADS[workload]
    label = workload + (" (bf16 decoder)" if dec_precision == "bf16" and precision != "bf16" else "")
    return {"workload": label, "model": cfg_name.lower(), "batch_per_gpu": B, "frames": T,
            "precision": "encoder GEMMs: " + PRECISION_TEXT[precision] + "; decoder/refiner GEMMs: "
                         + PRECISION_TEXT[dec_precision] + "; fp32 accumulate everywhere",
            "weights": "random-init (seed 0) + q_in_proj recalibration",
            "l2": "inputs and intermediates exceed L2 (134 MB mels, GBs of activations per step)",
            "parallelism": f"utterance shards x{world}, no collective"}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm
# ---------------------------------------------------------------------------------------------------------------
class CpuArm:
    """The reference's CPU implementation of the path: the UNMODIFIED reference from baseline/_ref (kind
    "reference"), or the oracle port when no copy of the reference is reachable (kind "port")."""

    def __init__(self, cfg, sd):
        from oracle import reference_runner as RR
        self.cfg, self.sd = cfg, sd
        self.threads = os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.ref = None
        try:
            self.ref = RR.import_preencoder()
        except Exception as e:                      # pragma: no cover - a broken copy falls back to the port, loudly
            print(f"bench: reference import failed ({e!r}); timing the oracle port instead", file=sys.stderr)
        if self.ref is not None:
            self.RR = RR
            self.model = RR.build_model(self.ref, cfg, sd)
            self.kind = "reference"
            self.what = f"unmodified reference PreEncoder.encode + decode ({RR.find()}), eager fp32, torch {torch.__version__} CPU"
        else:
            from oracle import preencoder_oracle as O
            self.O = O
            self.w = O.effective_weights(sd)
            self.kind = "port"
            self.what = f"oracle/preencoder_oracle.py (no copy of the reference reachable), torch {torch.__version__} CPU"

    def run(self, mel, lengths, micro_batch=4):
        """-> (indices, re-encoded mels, seconds).  Micro-batches of 4 utterances bound the reference's (B,C,C,T)
        expansion (1 MiB per frame per tensor) to ~8 GB; full-length utterances are batch-invariant."""
        t0 = time.perf_counter()
        if self.ref is not None:
            idx, out = self.RR.reencode(self.ref, self.model, mel, lengths, micro_batch=micro_batch)
        else:
            O = self.O
            mask = O.sequence_mask(mel.shape[1], lengths).unsqueeze(1)
            with torch.no_grad():
                idx = O.encode(self.w, self.cfg, mel, mask, folded=True)
                out = O.decode(self.w, self.cfg, idx, mask, folded=True)
        return idx, out, time.perf_counter() - t0

    def probe(self, T):
        """Seconds for one utterance of T frames (also the warm-up of the thread pool / allocator)."""
        from mqgan_b200.synth import synth_mels
        mel = synth_mels(1, T, self.cfg.mel_channels, seed=99)
        return self.run(mel, torch.full((1,), T, dtype=torch.long))[2]


def reference_arm(args):
    """--impl reference: each step = the reference's encode + decode on a bounded sample of the workload (utterances
    of configs[0]'s length, as many as keep K + W steps within a few minutes), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mqgan_b200 import spec as S
    from mqgan_b200.synth import synth_mels
    cfg_name, B, T, _, _ = WORKLOADS[args.workload]
    cfg = getattr(S, cfg_name)
    arm = CpuArm(cfg, calibrated_state_dict(cfg))
    st = min(CPU_CONFIG0[1], T)
    t1 = arm.probe(st)
    total_steps = args.steps + args.warmup
    sb = 1
    while sb < CPU_CONFIG0[0] and 2 * sb * t1 * total_steps <= args.cpu_budget_s:
        sb *= 2
    mel = synth_mels(sb, st, cfg.mel_channels, seed=1)
    lengths = torch.full((sb,), st, dtype=torch.long)
    for _ in range(args.warmup):
        arm.run(mel, lengths)
    times = [arm.run(mel, lengths)[2] for _ in range(args.steps)]
    ms = 1e3 * sum(times) / len(times)
    fps = sb * st / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_dict(args.workload, args.gpus),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": arm.threads, "kind": arm.kind,
                         "sample": f"each step: {sb} utterances x {st} frames (BASELINE configs[0] is 16 x 512; the count is "
                                   f"cut so that {total_steps} steps fit ~{args.cpu_budget_s:.0f} s at {t1:.1f} s per utterance), {arm.what}"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# GPU arm helpers
# ---------------------------------------------------------------------------------------------------------------
class Runner:
    """Timing plumbing shared by the headline and the secondary workloads."""

    def __init__(self, dev, dist, rank, world):
        self.dev, self.dist, self.rank, self.world = dev, dist, rank, world

    def barrier(self):
        torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """ms for ``steps`` calls: CUDA events on the launching stream, barrier + sync on both sides, max over ranks."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.dist is not None:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms


def time_reencode(run: Runner, model, cfg, B, T, steps, warmup, seed, sample_clocks=False):
    """Resident and end-to-end timing of encode + decode on one workload.  Returns a dict and keeps nothing alive."""
    from mqgan_b200 import _lib
    from mqgan_b200.synth import synth_mels
    dev = run.dev
    eng = model.engine()
    mel_host = synth_mels(B, T, cfg.mel_channels, seed=seed).pin_memory()
    out_host = torch.empty(B, T, cfg.mel_channels, dtype=torch.float32).pin_memory()
    idx_host = torch.empty(B, T, dtype=torch.int64).pin_memory()
    mel_dev = mel_host.to(dev, non_blocking=True)

    def step_resident():
        idx = eng.encode(mel_dev, None)
        return idx, eng.decode(idx, None)

    def step_e2e():
        x = mel_host.to(dev, non_blocking=True)
        idx = model.encode(x, None)
        idx_host.copy_(idx, non_blocking=True)
        model.decode(idx, None, host_out=out_host)        # chunk-wise D2H overlapped with the remaining chunks' compute

    for _ in range(warmup):
        step_resident()
    sampler = ClockSampler(dev.index or 0)
    if sample_clocks and run.rank == 0:
        sampler.start()
    l0 = _lib.launch_count
    ms_total = run.timed(step_resident, steps)
    launches = _lib.launch_count - l0
    clocks = sampler.stop() if (sample_clocks and run.rank == 0) else None
    step_e2e()
    ms_e2e = run.timed(step_e2e, steps)
    # per-kernel table: one instrumented step, CUDA events around every launch on the launching stream
    _lib.profiler = _lib.LaunchProfiler()
    step_resident()
    table = _lib.profiler.summary()
    _lib.profiler = None
    return {"ms_step": ms_total / steps, "ms_e2e": ms_e2e / steps, "launches": launches, "clocks": clocks, "table": table,
            "mel_host": mel_host, "h2d": mel_host.numel() * 4, "d2h": out_host.numel() * 4 + idx_host.numel() * 8}


def conv_roofline(table, peaks, cfg_name):
    conv = [r for r in table if r[0] == "mq_conv_gemm"]
    conv_ms = sum(r[3] for r in conv)
    conv_flops = sum(r[4] for r in conv)
    conv_mma = sum(r[5] for r in conv)
    n = sum(r[2] for r in conv)
    total_ms = sum(r[3] for r in table)
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    issued = conv_mma / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    peak = peaks["bf16_tflops_sustained"]
    roof = {
        "bound": "tensor",
        "kernel": "mq_conv_gemm (tcgen05 implicit-GEMM conv: conv_pair_kernel cta_group::2 for the refiner's 3x3 layers, "
                  "conv_pair1d_kernel / conv_gemm_kernel for 1-D / 1x1 layers; all %d launches of one step)" % n,
        # issued = MMA FLOPs the tensor pipe really executed (the fused up-conv pre-sums 9 -> 6 taps, so it issues fewer
        # than the reference's convolution counts; the fp32-grade split layers issue three products per algorithmic one)
        "issued_mma_tflops": issued, "frac_issued": issued / peak,
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "peak_source": peaks["source"] + " bf16 sustained",
        "traffic": None, "share_of_step": conv_ms / total_ms if total_ms > 0 else None,
        "flops_per_launch_avg": conv_flops / max(n, 1), "ms_per_launch_avg": conv_ms / max(n, 1),
        "kernel_ms_sum": total_ms,
    }
    # DRAM bytes per launch from the committed ncu capture of one full step (tools/ncu_traffic.py -> profiles/)
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")
    if os.path.exists(tpath):
        try:
            t = json.load(open(tpath)).get(cfg_name.lower(), {}).get("mq_conv_gemm")
            if t:
                roof["traffic"] = t["dram_bytes_per_launch"]
                roof["traffic_algorithmic_bytes"] = t.get("algorithmic_bytes_per_launch")
                roof["traffic_source"] = t.get("source")
        except Exception:
            pass
    return roof


def write_layer_table(path, table, workload, frames):
    total_ms = sum(r[3] for r in table)
    rows = sorted(table, key=lambda r: -r[3])
    with open(path, "w") as f:
        f.write(f"# per-kernel table, workload {workload}, one step, CUDA events on the launching stream\n\n")
        f.write("| entry point | layer | launches | ms | share | algorithmic TFLOP/s | issued MMA TFLOP/s |\n|---|---|---|---|---|---|---|\n")
        for name, tag, n, ms, fl, mma in rows:
            tf = f"{fl / (ms * 1e-3) / 1e12:.1f}" if fl > 0 and ms > 0 else ""
            tm = f"{mma / (ms * 1e-3) / 1e12:.1f}" if mma > 0 and ms > 0 else ""
            f.write(f"| {name} | {tag} | {n} | {ms:.3f} | {100 * ms / total_ms:.1f}% | {tf} | {tm} |\n")
        f.write(f"\ntotal {total_ms:.3f} ms for {frames} frames\n")


def reference_on_gpu(cfg, sd, dev):
    """The UNMODIFIED reference PreEncoder as an fp32 CUDA module (TF32 off) - a second checker for the index gate:
    north_star asks for indices "bit-exact against the fp32 reference", and this is that reference, run on the same
    box over the same frames.  None when no copy of the reference is reachable."""
    try:
        from oracle import reference_runner as RR
        ref = RR.import_preencoder()
        if ref is None:
            return None
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        return RR.build_model(ref, cfg, sd).to(dev)
    except Exception as e:                                       # pragma: no cover
        print(f"bench: reference-on-GPU checker unavailable ({e!r})", file=sys.stderr)
        return None


def full_parity(model, sd, cfg, mel_host, dev, mel_frames=16384, chunk=16, ref_model=None):
    """Checker, outside every timed region: float64 restatement on the GPU (oracle/gpu_checker.py) over ALL frames of
    the workload for the indices, and over ``mel_utts`` utterances for the re-encoded mels."""
    from oracle import gpu_checker as G
    B, T, _ = mel_host.shape
    mel_utts = max(1, min(B, mel_frames // T))
    t0 = time.perf_counter()
    w64 = G.weights_on(sd, dev, torch.float64)
    mism = safe_mism = safe = 0
    tau = 2e-4
    mel_err = mel_ref_max = 0.0
    num = den = 0.0
    ref32 = {"frames": 0, "mismatch": 0, "mismatch_outside_tau": 0, "ref32_vs_fp64_mismatch": 0}
    with torch.no_grad():
        idx_gpu = model.encode(mel_host.to(dev), None)
        per = max(1, min(chunk, (16 * 1024) // T if T > 1024 else chunk))
        for b0 in range(0, B, per):
            z = G.encode_latents(w64, cfg, mel_host[b0:b0 + per].to(dev))
            ref_idx, margin = G.fsq_indices_and_margin(z, cfg.fsq_levels)
            neq = ref_idx != idx_gpu[b0:b0 + per]
            ok = margin > tau
            mism += int(neq.sum())
            safe += int(ok.sum())
            safe_mism += int((neq & ok).sum())
            if ref_model is not None:                    # the reference's own fp32 indices (micro-batches of <= 4096 frames:
                sub_r = max(1, 4096 // T)                # its ConvBlock2D expands to 1 MiB per frame per tensor)
                for c0 in range(0, min(per, B - b0), sub_r):
                    xr = mel_host[b0 + c0:b0 + c0 + sub_r].to(dev)
                    ir = ref_model.encode(xr, None)
                    ng = ir != idx_gpu[b0 + c0:b0 + c0 + sub_r]
                    ref32["mismatch"] += int(ng.sum())
                    ref32["mismatch_outside_tau"] += int((ng & ok[c0:c0 + sub_r]).sum())
                    ref32["ref32_vs_fp64_mismatch"] += int((ir != ref_idx[c0:c0 + sub_r]).sum())
                    ref32["frames"] += int(ir.numel())
            if b0 < mel_utts:
                n = min(per, mel_utts - b0)
                sub = max(1, 4096 // T)              # the float64 refiner holds ~1 MB per frame
                for c0 in range(0, n, sub):
                    sl = slice(b0 + c0, b0 + min(c0 + sub, n))
                    ref = G.decode(w64, cfg, ref_idx[sl.start - b0:sl.stop - b0])
                    out = model.decode(ref_idx[sl.start - b0:sl.stop - b0], None).double()
                    d = out - ref
                    mel_err = max(mel_err, float(d.abs().max()))
                    mel_ref_max = max(mel_ref_max, float(ref.abs().max()))
                    num += float((d * d).sum())
                    den += float((ref * ref).sum())
    torch.cuda.synchronize()
    frames = B * T
    extra = {}
    if ref_model is not None and ref32["frames"] > 0:
        ref32["index_match"] = 1.0 - ref32["mismatch"] / ref32["frames"]
        ref32["what"] = ("the unmodified reference PreEncoder.encode as an fp32 CUDA module (allow_tf32 off) on the same frames; "
                         "ref32_vs_fp64_mismatch = how often the reference's own fp32 answer differs from the float64 checker")
        extra["vs_reference_fp32_gpu"] = ref32
    return {**extra, "checker": "float64 restatement on the GPU (oracle/gpu_checker.py, pinned to the oracle / reference goldens in "
                       "tests/test_oracle_golden.py); outside the timed region",
            "frames": frames, "index_mismatches": mism, "index_match": 1.0 - mism / frames,
            "safe_frames": safe, "safe_mismatch": safe_mism, "margin_tau": tau,
            "mel_utterances": min(mel_utts, B), "mel_max_abs_err": mel_err, "mel_ref_max_abs": mel_ref_max,
            "mel_rel_l2": (num / den) ** 0.5 if den > 0 else None, "checker_seconds": time.perf_counter() - t0}


def vq_microbench(dev, peaks, n=1 << 20):
    """BASELINE configs[3]: N = 2^20 latents against K in {1024, 8192} codes; D = 4 / 5 (the reference's only quantiser
    widths, FSQ implicit codebooks) and D = 64 (random normal codebook).  Roofline per case: HBM for D <= 5 (algorithmic
    bytes = z in + idx + codes out), tensor (2NKD) for D = 64, and for every case the TMEM-read ceiling of a kernel
    that must read each of the N x K fp32 scores from tensor memory once (64 B/clk/SM)."""
    import numpy as np
    from mqgan_b200 import ops, _lib

    def implicit_codebook(levels):
        """FSQ's implicit codebook (quantizer.py:101-104, 183-187): every index's code in [-1, 1]^D - input data of the
        microbench, built here (the oracle is only ever the checker)."""
        lv = torch.tensor(levels)
        basis = torch.cumprod(torch.tensor([1] + list(levels[:-1])), dim=0)
        digits = (torch.arange(int(np.prod(levels)))[:, None] // basis) % lv
        hw = lv // 2
        return ((digits - hw) / hw).float()
    # measured ceiling: the chip's TMEM read bandwidth with the epilogue's own tcgen05.ld pattern and nothing else
    sink = torch.zeros(1, device=dev)
    iters = 2000
    stream = torch.cuda.current_stream().cuda_stream
    _lib.call("mq_tmem_read_probe", 50, sink.data_ptr(), stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call("mq_tmem_read_probe", iters, sink.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    sms = _lib.lib().mq_sm_count()
    tmem_gbs = sms * iters * 128 * 512 * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    out = [{"tmem_read_probe": {"gbs_chip": tmem_gbs, "gbs_per_sm": tmem_gbs / sms, "how": "mq_tmem_read_probe: 8 warps per SM reading "
            "128 lanes x 512 columns with tcgen05.ld.32x32b.x32, no MMA, no math; CUDA events"}}]
    cases = [("fsq[8,8,4,4]", [8, 8, 4, 4], None), ("fsq[8,8,8,4,4]", [8, 8, 8, 4, 4], None),
             ("random K=1024 D=64", None, (1024, 64)), ("random K=8192 D=64", None, (8192, 64))]
    for name, levels, shape in cases:
        g = torch.Generator().manual_seed(0)
        if levels is not None:
            K, D = int(np.prod(levels)), len(levels)
            cb = implicit_codebook(levels)
            z = (torch.randn(n, D, generator=g) * 0.6).clamp(-1.05, 1.05)
        else:
            K, D = shape
            cb = torch.randn(K, D, generator=g)
            z = torch.randn(n, D, generator=g)
        zd = z.to(dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for prec in ("f16x2", "bf16"):
            pc = ops.pack_codebook(cb, prec).to(dev)
            for _ in range(3):
                idx, codes = ops.vq_nearest(zd, pc)
            reps, ms = 10, 0.0
            for _ in range(reps):
                flush.fill_(1)                                  # L2 flush between timed launches (48 MB working set < 126 MB L2)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                idx, codes = ops.vq_nearest(zd, pc)
                e1.record()
                torch.cuda.synchronize()
                ms += e0.elapsed_time(e1)
            ms /= reps
            ref = torch.argmin(torch.cdist(zd[:65536].double(), cb.to(dev).double()), dim=1)
            agree = float((idx[:65536] == ref).float().mean())
            bytes_alg = n * D * 4 + n * (8 + D * 4)
            rec = {"case": name, "precision": prec, "n": n, "k": K, "d": D, "ms": ms,
                   "index_agreement_vs_fp64_argmin_64k_rows": agree}
            k_pad = pc.k_pad
            if D <= 16:
                rec["roofline"] = {"bound": "hbm", "achieved": bytes_alg / ms / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": bytes_alg / ms / 1e6 / peaks["hbm_gbs"], "algorithmic_bytes": bytes_alg}
            else:
                fl = 2.0 * n * K * D
                rec["roofline"] = {"bound": "tensor", "achieved": fl / ms / 1e9, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                   "frac": fl / ms / 1e9 / peaks["bf16_tflops"], "algorithmic_flops": fl,
                                   "issued_mma_tflops": fl * (3 if prec == "f16x2" else 1) / ms / 1e9}
            # every score is one 4-byte TMEM word that has to be read once: n * k_pad * 4 bytes at the measured TMEM rate
            tmem_ms = n * k_pad * 4 / (tmem_gbs * 1e9) * 1e3
            rec["tmem_read_ceiling"] = {"ms": tmem_ms, "frac": tmem_ms / ms, "score_bytes": n * k_pad * 4}
            out.append(rec)
        del flush
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=None, choices=["f16x2", "bf16x3", "bf16"],
                    help="override the workload's encoder operand format")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="headline only (skip the other BASELINE configs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-size float64 parity check")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0, help="--impl reference: wall-clock target for all steps")
    ap.add_argument("--secondary-timeout-s", type=float, default=900.0, help="watchdog for the secondary block")
    ap.add_argument("--layer-table", default=None, help="write the per-layer kernel table (markdown) here")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return
    if args.warmup < 3:
        args.warmup = 3

    from mqgan_b200 import spec as S, _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.lib().mq_device_check(), "mq_device_check")
    run = Runner(dev, dist, rank, world)
    peaks = load_peaks()

    cfg_name, B, T, precision, dec_precision = WORKLOADS[args.workload]
    precision = args.precision or precision
    cfg = getattr(S, cfg_name)
    model, sd = build_model(cfg, precision, dev, decoder_precision=dec_precision)
    frames = B * T
    head = time_reencode(run, model, cfg, B, T, args.steps, args.warmup, seed=1 + rank, sample_clocks=True)
    roofline = conv_roofline(head["table"], peaks, cfg_name)
    if rank == 0 and args.layer_table:
        write_layer_table(args.layer_table, head["table"], args.workload, frames)
    from mqgan_b200.spec import flops_per_frame
    fpf = flops_per_frame(cfg)["total"]
    ms_step = head["ms_step"]
    value = world * frames / (ms_step * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if dec_precision == "bf16" else dec_precision, "data": "synthetic",
        "config": config_dict(args.workload, world),
        "e2e": {"value": world * frames / (head["ms_e2e"] * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": head["h2d"],
                "d2h_bytes_per_step": head["d2h"], "ms_per_step": head["ms_e2e"]},
        "gpu_launches": head["launches"], "clocks": head["clocks"], "roofline": roofline,
        # timed step minus the sum of per-launch CUDA-event durations of one instrumented step: <= 0 means the launch queue
        # never runs dry (the host enqueues ~200 launches per step far ahead of the GPU); event pairs around single
        # launches over-count by the inter-kernel gap, hence slightly negative values
        "step_ms_minus_kernel_ms_sum": ms_step - roofline["kernel_ms_sum"],
        "pass_frac_of_tensor_roofline": value / world * fpf / (peaks["bf16_tflops_sustained"] * 1e12),
        "cpu_baseline": None, "parity": None, "secondary": None,
    }

    # ---------------- parity at workload size + CPU baseline (rank 0, N = 1 only; never timed) ----------------
    mel_host = head.pop("mel_host")
    if world == 1 and not args.no_parity:
        try:
            tf32_flags = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
            ref_gpu = reference_on_gpu(cfg, sd, dev)
            line["parity"] = full_parity(model, sd, cfg, mel_host, dev, ref_model=ref_gpu)
            del ref_gpu
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32_flags
        except Exception as e:                                   # the headline must survive a checker failure
            line["parity"] = {"error": repr(e)}
    if world == 1 and not args.no_cpu_baseline:
        try:
            arm = CpuArm(cfg, sd)
            st = min(CPU_CONFIG0[1], T)
            t1 = arm.probe(st)
            sb = CPU_CONFIG0[0] if CPU_CONFIG0[0] * t1 <= 60.0 else max(1, int(30.0 / t1))
            sub = mel_host[:sb, :st].clone()
            lengths = torch.full((sb,), st, dtype=torch.long)
            ref_idx, ref_out, dt = arm.run(sub, lengths)
            line["cpu_baseline"] = {"value": sb * st / dt, "unit": "frames/s", "cores": arm.threads, "kind": arm.kind,
                                    "sample": f"BASELINE configs[0] shape: first {sb} utterances x first {st} frames of the workload batch, "
                                              f"one run of {dt:.1f} s in micro-batches of 4, {arm.what}"}
            g_idx = model.encode(sub.to(dev), None).cpu()
            g_out = model.decode(ref_idx.to(dev), None).cpu()
            d = g_out - ref_out
            if isinstance(line["parity"], dict):
                line["parity"]["vs_cpu_arm"] = {
                    "kind": arm.kind, "frames": sb * st, "index_mismatches": int((g_idx != ref_idx).sum()),
                    "index_match": float((g_idx == ref_idx).float().mean()), "mel_max_abs_err": float(d.abs().max()),
                    "mel_ref_max_abs": float(ref_out.abs().max()),
                    "mel_rel_l2": float((d.double().pow(2).sum() / ref_out.double().pow(2).sum()).sqrt())}
            del arm
        except Exception as e:
            line["cpu_baseline"] = {"error": repr(e)}
    del model, mel_host
    torch.cuda.empty_cache()

    # ---------------- the other BASELINE configs, each timed like the headline ----------------
    def emit(tag=None):
        if tag is not None:
            line["secondary_incomplete"] = tag
        if rank == 0:
            print(json.dumps(line), flush=True)

    if not args.no_secondary:
        sec = {}
        line["secondary"] = sec
        # safety net: the secondary block must never cost the headline line.  If it has not finished in time (a hung
        # collective on some rank, a stuck capture), every rank's timer fires, rank 0 prints what has been measured and
        # the processes leave without touching NCCL again.
        def _bail():
            emit(f"secondary block exceeded {args.secondary_timeout_s:.0f} s; printed what was measured")
            sys.stdout.flush()
            os._exit(0)
        watchdog = threading.Timer(args.secondary_timeout_s, _bail)
        watchdog.daemon = True
        watchdog.start()
        ssteps = max(3, min(args.steps, 5))

        def reencode_line(name, want_parity):
            c_name, b, t, prec, dprec = WORKLOADS[name]
            c = getattr(S, c_name)
            m, sd2 = build_model(c, prec, dev, decoder_precision=dprec)
            r = time_reencode(run, m, c, b, t, ssteps, 3, seed=11 + rank)
            roof = conv_roofline(r["table"], peaks, c_name)
            v = world * b * t / (r["ms_step"] * 1e-3)
            ops_per_alg = 3.0 if dprec == "f16x2" else 1.0
            o = {"metric": METRIC, "value": v, "unit": "frames/s", "ms_per_step": r["ms_step"], "steps": ssteps, "warmup": 3,
                 "dtype": dprec, "config": config_dict(name, world),
                 "e2e": {"value": world * b * t / (r["ms_e2e"] * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": r["h2d"],
                         "d2h_bytes_per_step": r["d2h"]},
                 "gpu_launches": r["launches"], "roofline": roof,
                 "pass_frac_of_tensor_roofline": v / world * flops_per_frame(c)["total"] / (peaks["bf16_tflops_sustained"] * 1e12),
                 "pass_frac_issued": v / world * flops_per_frame(c)["total"] * ops_per_alg / (peaks["bf16_tflops_sustained"] * 1e12)}
            if want_parity and world == 1 and not args.no_parity:
                try:
                    o["parity"] = full_parity(m, sd2, c, r["mel_host"], dev)
                except Exception as e:
                    o["parity"] = {"error": repr(e)}
            del m, r
            torch.cuda.empty_cache()
            return o

        for key, name, par in (("hifispeech_fp32_decoder", "hifispeech_256x1024_fp32", True),
                               ("hifimusic_bf16", "hifimusic_32x8192_bf16", True)):
            if name == args.workload:
                continue
            try:
                from mqgan_b200.engine import PreEncoderEngine
                if WORKLOADS[name][4] != "bf16" and "decoder_precision" not in PreEncoderEngine.__init__.__code__.co_varnames:
                    raise NotImplementedError("decoder_precision mode not built")
                sec[key] = reencode_line(name, par)
            except Exception as e:
                sec[key] = {"error": repr(e)}
                if world > 1:
                    raise
        if world == 1:
            try:
                sec["vq_lookup"] = vq_microbench(dev, peaks)
            except Exception as e:
                sec["vq_lookup"] = {"error": repr(e)}
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import train_bench as TB
            targs = TB.parse_args(["--steps", str(max(5, min(args.steps, 10))), "--warmup", "3"])
            tl = TB.measure(targs)
            if tl is not None:
                sec["train_step"] = {k: tl[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "dtype", "config",
                                                        "e2e", "gpu_launches", "roofline", "native_share_of_step",
                                                        "step_tflops_algorithmic", "replicas_in_sync", "scaling", "n_gpus")}
        except Exception as e:
            sec["train_step"] = {"error": repr(e)}
            if world > 1:
                raise
        if world == 1 and "error" not in sec.get("train_step", {}):
            # the same step with the discriminator convolutions on the library's tcgen05 kernels too (TrainStep(d_native=True)):
            # slower than cuDNN's strided kernels, so not the default - reported for its native share
            try:
                targs = TB.parse_args(["--steps", "5", "--warmup", "3", "--d_native"])
                tl = TB.measure(targs)
                if tl is not None:
                    sec["train_step_d_native"] = {k: tl[k] for k in ("value", "unit", "ms_per_step", "native_share_of_step",
                                                                     "gpu_launches", "config")}
            except Exception as e:
                sec["train_step_d_native"] = {"error": repr(e)}
        watchdog.cancel()

    emit()
    if dist is not None:
        # the training step's captured graphs hold NCCL work: tearing the communicator down under them hangs
        sys.stdout.flush()
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    main()
