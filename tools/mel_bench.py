"""Micro-benchmark of mq_log_mel: B utterances x S seconds at the hifispeech spec.
Prints frames/s, algorithmic HBM bytes (4 B per sample in, n_mels * 4 B per frame out) and GB/s against
the measured copy bandwidth, and the CPU oracle's frames/s on one utterance."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import yaml
from mqgan_b200.melspec import LogMelExtractor
from oracle import mel_oracle as M

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
SEC = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
spec = {"sampling_rate": 44100, "filter_length": 2048, "hop_length": 512, "win_length": 2048,
        "n_mel_channels": 128, "mel_fmin": 0.0, "mel_fmax": 22050.0}
n = int(SEC * spec["sampling_rate"])
ext = LogMelExtractor(spec, "cuda")
wav = (torch.rand(B, n, device="cuda") - 0.5)
for _ in range(3):
    out, frames = ext(wav)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
reps = 10
for _ in range(reps):
    out, frames = ext(wav)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
nfr = sum(frames)
bytes_alg = B * n * 4 + nfr * spec["n_mel_channels"] * 4
peak = 6534.8
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
w1 = wav[0].cpu().numpy()
t0 = time.perf_counter()
ref = M.log_mel(w1, spec, np.float32)
cpu_s = time.perf_counter() - t0
err = float(np.abs(out[0, : frames[0]].cpu().numpy() - M.log_mel(w1, spec, np.float64)).max())
print(json.dumps({"kernel": "mq_log_mel", "utterances": B, "seconds_each": SEC, "frames": nfr, "ms": ms,
                  "frames_per_s": nfr / ms * 1e3, "audio_seconds_per_s": B * SEC / ms * 1e3,
                  "alg_bytes": bytes_alg, "gbs": bytes_alg / ms / 1e6, "frac_of_hbm_peak": bytes_alg / ms / 1e6 / peak,
                  "cpu_oracle_frames_per_s_1core": frames[0] / cpu_s, "max_abs_err_vs_float64": err}))
