"""Where one eager training iteration spends its GPU time, by phase (CUDA events at the phase boundaries of
TrainStep._step_body) and by library vs non-library kernels inside each phase.
Usage: python tools/train_phases.py [out.md]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mqgan_b200 import _lib, spec as S
from mqgan_b200 import training as TR
from mqgan_b200.synth import synth_disc_state_dict, synth_mels, synth_state_dict

cfg, pdc, mbc = S.HIFISPEECH, S.HIFISPEECH_PATCH_D, S.HIFISPEECH_MULTIBIN_D
B, T = 16, 256
ts = TR.TrainStep(cfg, pdc, mbc, synth_state_dict(cfg, 0), synth_disc_state_dict(S.patch_disc_param_spec(pdc), 1),
                  synth_disc_state_dict(S.multibin_param_spec(mbc), 2), dict(S.TRAIN_DEFAULTS), "cuda", d_autocast_bf16=True)
x = synth_mels(B, T, cfg.mel_channels, seed=1).cuda()
lens = torch.full((B,), T, dtype=torch.long, device="cuda")
for _ in range(3):
    ts.step(x, lens)
torch.cuda.synchronize()
ts.phase_events = []
_lib.profiler = _lib.LaunchProfiler()
ts.step(x, lens)
torch.cuda.synchronize()
recs = _lib.profiler.records
_lib.profiler = None
marks = ts.phase_events
ts.phase_events = None
lines = ["# one eager training iteration (hifispeech 16 x 256), GPU time by phase\n",
         "Eager launches serialise on the host, so phase times are upper bounds of what the CUDA-graph replay (45.6 ms for the",
         "whole iteration) spends; the split between the library's kernels and everything else is what matters.\n",
         "| phase | ms (eager) | library kernels ms | library launches | other (PyTorch / cuDNN) ms |", "|---|---|---|---|---|"]
tot = lib_tot = 0.0
for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
    ms = e0.elapsed_time(e1)
    lib_ms, cnt = 0.0, 0
    for name, meta, s0, s1 in recs:
        if e0.elapsed_time(s0) >= 0 and s1.elapsed_time(e1) >= 0:
            lib_ms += s0.elapsed_time(s1)
            cnt += 1
    tot += ms
    lib_tot += lib_ms
    lines.append(f"| {n1} | {ms:.2f} | {lib_ms:.2f} | {cnt} | {ms - lib_ms:.2f} |")
lines.append(f"| total | {tot:.2f} | {lib_tot:.2f} | {len(recs)} | {tot - lib_tot:.2f} |")
out = "\n".join(lines)
print(out)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(out + "\n")
