"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np
import torch

from mqgan_b200 import spec as S
from mqgan_b200.synth import synth_state_dict, synth_mels, amplify_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    """Returns (cfg, state_dict with the fixture's calibrated q_in_proj, mel, lengths, fixture)."""
    fx = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = getattr(S, str(fx["config"]))
    sd = synth_state_dict(cfg, seed=int(fx["seed"]))
    if "amplified" in fx.files and int(fx["amplified"]):
        sd = amplify_state_dict(sd)
    sd["q_in_proj.weight"] = torch.from_numpy(fx["qin_w"]).clone()
    sd["q_in_proj.bias"] = torch.from_numpy(fx["qin_b"]).clone()
    lengths = torch.from_numpy(fx["lengths"]).long()
    T = int(fx["T"])
    mel = synth_mels(len(lengths), T, cfg.mel_channels, seed=int(fx["mel_seed"]))
    pad = torch.arange(T)[None, :] >= lengths[:, None]
    mel = mel.masked_fill(pad.unsqueeze(-1), 0.0)
    return cfg, sd, mel, lengths, fx


def index_report(idx, ref_idx, margin=None, tau=1e-4):
    """Raw agreement and margin-aware agreement: frames whose fp64
    distance to a rounding boundary exceeds tau must agree."""
    idx = torch.as_tensor(idx).long().cpu()
    ref_idx = torch.as_tensor(ref_idx).long().cpu()
    neq = idx != ref_idx
    rep = {"frames": int(idx.numel()), "mismatch": int(neq.sum()),
           "agree": 1.0 - float(neq.float().mean())}
    if margin is not None:
        safe = margin.cpu() > tau
        rep["safe_frames"] = int(safe.sum())
        rep["safe_mismatch"] = int((neq & safe).sum())
    return rep
