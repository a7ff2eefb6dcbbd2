"""torch.profiler breakdown of one training step (development aid): top CUDA kernels and CPU ops.
Usage: python tools/train_profile.py [rows]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from mqgan_b200 import spec as S
from mqgan_b200 import training as TR
from mqgan_b200.synth import synth_disc_state_dict, synth_mels, synth_state_dict

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 40
cfg, pdc, mbc = S.HIFISPEECH, S.HIFISPEECH_PATCH_D, S.HIFISPEECH_MULTIBIN_D
B, T = 16, 256
ts = TR.TrainStep(cfg, pdc, mbc, synth_state_dict(cfg, 0), synth_disc_state_dict(S.patch_disc_param_spec(pdc), 1),
                  synth_disc_state_dict(S.multibin_param_spec(mbc), 2), dict(S.TRAIN_DEFAULTS), "cuda", d_autocast_bf16=True)
x = synth_mels(B, T, cfg.mel_channels, seed=1).cuda()
lens = torch.full((B,), T, dtype=torch.long, device="cuda")
for _ in range(3):
    ts.step(x, lens)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    ts.step(x, lens)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=rows, max_name_column_width=70))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=70))
