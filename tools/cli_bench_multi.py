"""The sharded re-encode CLI at 1 / 2 / 4 / 8 GPUs, including .npy file I/O (reference path
reencode_spectrograms_from_checkpoint.py:60-105 with the additive --gpus flag).

Writes n_files synthetic log-mel files (300 - 1100 frames, the reference's (T, n_mels) float32 layout) under a tmpfs
directory, saves a random-init hifispeech checkpoint + config, and runs the real CLI script once per GPU count.  Each
worker reports its model-load time and its processing window; frames/s = all frames / (latest end - earliest start), so
the per-process start-up (import, weight load, table build) is reported separately from the steady-state rate.

    python tools/cli_bench_multi.py [n_files] [gpu counts, e.g. 1,2,4,8] [batch_size] [--sort]
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import yaml

from mqgan_b200 import spec as S
from mqgan_b200.synth import synth_state_dict

n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
counts = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8").split(",")]
batch_size = int(sys.argv[3]) if len(sys.argv) > 3 else 32
sort = "--sort" in sys.argv
base = os.environ.get("MQ_BENCH_BASE") or ("/dev/shm" if os.path.isdir("/dev/shm") else None)
keep = os.environ.get("MQ_BENCH_TREE")            # reuse (and keep) an input tree across invocations
root = keep or tempfile.mkdtemp(prefix="mq_cli_multi_", dir=base)
cfg = S.HIFISPEECH
src = os.path.join(root, "in")
rng = np.random.default_rng(0)
frames = 0
t0 = time.perf_counter()
fresh = not os.path.isdir(src)
for i in range(n_files):
    d = os.path.join(src, f"spk{i % 64:02d}")
    T = int(rng.integers(300, 1100))
    frames += T
    if fresh:
        os.makedirs(d, exist_ok=True)
        np.save(os.path.join(d, f"u{i:06d}.npy"), (rng.standard_normal((T, cfg.mel_channels)) * 2 - 4).astype(np.float32))
if not fresh:                                     # reused tree: count the frames from the files themselves
    frames = 0
    for dp, _, fs in os.walk(src):
        for fn in fs:
            frames += int(np.load(os.path.join(dp, fn), mmap_mode="r").shape[0])
t_write = time.perf_counter() - t0
ckpt = os.path.join(root, "ckpt.pth")
torch.save({"model_state_dict": synth_state_dict(cfg, 0)}, ckpt)
conf = os.path.join(root, "config.yaml")
with open(conf, "w") as f:
    yaml.safe_dump({"model": {"mel_channels": cfg.mel_channels, "generator": {
        "channels": list(cfg.channels), "kernel_sizes": list(cfg.kernel_sizes), "fsq_levels": list(cfg.fsq_levels),
        "refiner_base_channels": cfg.refiner_base_channels, "refiner_depth": cfg.refiner_depth,
        "refiner_hidden_proj_divisor": cfg.refiner_hidden_proj_divisor}}}, f)
env = dict(os.environ, MQ_CLI_TIMING="1")
for n in counts:
    dst = os.path.join(root, f"out{n}")
    cmd = [sys.executable, os.path.join(ROOT, "reencode_spectrograms_from_checkpoint.py"), "--checkpoint", ckpt, "--config", conf,
           "--input_dir", src, "--output_dir", dst, "--device", "cuda", "--batch_size", str(batch_size), "--gpus", str(n)]
    if sort:
        cmd.append("--sort_by_length")
    t0 = time.perf_counter()
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    wall = time.perf_counter() - t0
    timing = None
    for ln in res.stdout.splitlines():
        if ln.startswith("MQ_CLI_TIMING "):
            timing = json.loads(ln[len("MQ_CLI_TIMING "):])
    written = sum(len(fs) for _, _, fs in os.walk(dst)) if os.path.isdir(dst) else 0
    rec = {"tool": "cli_bench_multi", "gpus": n, "files": n_files, "files_written": written, "frames": frames, "batch_size": batch_size,
           "sort_by_length": sort, "wall_s_incl_startup": wall, "rc": res.returncode, "timing": timing,
           "frames_per_s": frames / timing["process_span_s"] if timing else None,
           "files_per_s": n_files / timing["process_span_s"] if timing else None,
           "io_gb_per_s": 2 * frames * cfg.mel_channels * 4 / 1e9 / timing["process_span_s"] if timing else None,
           "host_cores": os.cpu_count(), "tree": root, "setup_write_s": t_write,
           "env": {k: os.environ[k] for k in ("MQ_IO_THREADS", "OMP_NUM_THREADS", "CUDA_DEVICE_SCHEDULE", "MQ_BENCH_BASE", "MQ_WORKER_THREADS") if k in os.environ}}
    if res.returncode != 0:
        rec["stderr_tail"] = res.stderr[-800:]
    print(json.dumps(rec), flush=True)
    shutil.rmtree(dst, ignore_errors=True)
if not keep:
    shutil.rmtree(root, ignore_errors=True)
