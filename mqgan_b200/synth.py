"""Deterministic synthetic weights and mels for parity tests and the bench.

There is no network for checkpoints or datasets, so weights are random-init of
the reference architecture and mels are synthetic.  Everything
here is a pure function of (config, seed) using torch's CPU generator, so the
container that produced ``tests/golden/`` (by running the real reference) and
the GPU box regenerate the same tensors without shipping 120 MB of weights.

``recalibrate_q_in_proj`` is the SURVEY §0-D4 fix: with default init every
frame quantises to one index, which would make the index-parity gate vacuous;
an affine edit of ``q_in_proj`` makes the pre-quantiser latents zero-mean /
unit-std on a calibration batch.  It is applied to the *state-dict*, so the
reference / oracle and the CUDA path see identical weights.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Tuple

import torch

from .spec import PreEncoderConfig, param_spec


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synth_state_dict(cfg: PreEncoderConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init state-dict with the reference's key names and shapes.

    Weights / biases ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (PyTorch's default
    Linear/Conv scale); weight-norm ``g`` is set to ||v|| per output channel as
    ``weight_norm`` itself does at wrap time, so the effective weight equals
    ``v``; APTx beta = 1, gamma = 0.5 (attentions.py:17).
    """
    spec = param_spec(cfg)
    shapes = dict(spec)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape in spec:
        if key.endswith(".relu.beta"):
            sd[key] = torch.tensor(1.0)
            continue
        if key.endswith(".relu.gamma"):
            sd[key] = torch.tensor(0.5)
            continue
        if key.endswith("original0") or key.endswith("weight_g"):
            continue  # filled from v below
        if key.endswith(".bias"):
            base = key[: -len(".bias")]
            wkey = next(k for k in (base + ".weight", base + ".parametrizations.weight.original1",
                                    base + ".weight_v") if k in shapes)
            wshape = shapes[wkey]
        else:
            wshape = shape
        fan_in = 1
        for d in wshape[1:]:
            fan_in *= d
        bound = 1.0 / math.sqrt(max(fan_in, 1))
        t = (torch.rand(shape, generator=_gen(key, seed), dtype=torch.float32) * 2.0 - 1.0) * bound
        sd[key] = t
    for key, shape in spec:
        if key.endswith("original0"):
            v = sd[key[: -len("original0")] + "original1"]
        elif key.endswith("weight_g"):
            v = sd[key[: -len("weight_g")] + "weight_v"]
        else:
            continue
        sd[key] = v.reshape(v.shape[0], -1).norm(dim=1).reshape(shape).clone()
    return {k: sd[k] for k, _ in spec}


def synth_mels(batch: int, frames: int, n_mels: int, seed: int = 0) -> torch.Tensor:
    """mel = randn * 2 - 4: covers the log-mel range log(clamp(mag, 1e-5)) of
    convert_spectrograms.py:34."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * seed + 17)
    return torch.randn(batch, frames, n_mels, generator=g, dtype=torch.float32) * 2.0 - 4.0


def synth_lengths(batch: int, frames: int, seed: int = 0, ragged: bool = True) -> torch.Tensor:
    """Per-utterance lengths; ragged ~ U[frames/4, frames] with the first one full."""
    if not ragged:
        return torch.full((batch,), frames, dtype=torch.long)
    g = torch.Generator(device="cpu")
    g.manual_seed(7919 * seed + 3)
    lo = max(1, frames // 4)
    lens = torch.randint(lo, frames + 1, (batch,), generator=g, dtype=torch.long)
    lens[0] = frames
    return lens


def recalibrate_q_in_proj(sd: Dict[str, torch.Tensor], latents: torch.Tensor,
                          valid: torch.Tensor | None = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Rescale ``q_in_proj`` in place so that ``z = q_in_proj(y)`` is zero-mean /
    unit-std per dim on the calibration batch.

    ``latents`` is z computed with the *current* q_in_proj, shape (..., D).
    z' = (z - mu) / sigma  <=>  W' = W / sigma[:,None],  b' = (b - mu) / sigma.
    Returns (W', b').
    """
    z = latents.reshape(-1, latents.shape[-1]).double()
    if valid is not None:
        z = z[valid.reshape(-1)]
    mu = z.mean(dim=0)
    sigma = z.std(dim=0).clamp_min(1e-12)
    w = (sd["q_in_proj.weight"].double() / sigma[:, None]).float()
    b = ((sd["q_in_proj.bias"].double() - mu) / sigma).float()
    sd["q_in_proj.weight"] = w
    sd["q_in_proj.bias"] = b
    return w, b


def synth_disc_state_dict(spec, seed: int = 0, zero_bias: bool = False) -> Dict[str, torch.Tensor]:
    """Seeded state-dict for a discriminator from ``spec.patch_disc_param_spec`` / ``multibin_param_spec``:
    conv weights ~ N(0, 0.02) as discriminators.py:193-198 initialises them, small non-zero biases (the
    reference's zeros would leave the bias path untested), unit-norm spectral-norm vectors u and v, and
    PyTorch-default uniform squeeze-excite linears.  ``zero_bias=True`` gives the reference's own init (conv biases
    zero, discriminators.py:177-181): what the training CLI uses; the tests keep the non-zero variant."""
    sd: Dict[str, torch.Tensor] = {}
    shapes = dict(spec)
    for key, shape in spec:
        g = _gen("disc." + key, seed)
        if key.endswith(".weight_orig"):
            t = torch.randn(shape, generator=g) * 0.02
        elif key.endswith(".weight_u") or key.endswith(".weight_v"):
            t = torch.randn(shape, generator=g)
            t = t / t.norm().clamp_min(1e-12)
        elif key.endswith(".bias") and (key[: -len(".bias")] + ".weight_orig") in shapes:
            t = torch.randn(shape, generator=g) * 0.01
            if zero_bias:
                t = torch.zeros(shape)
        else:
            wshape = shapes[key[: -len(".bias")] + ".weight"] if key.endswith(".bias") else shape
            bound = 1.0 / math.sqrt(max(wshape[1], 1))
            t = (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound
        sd[key] = t
    return sd


def amplify_state_dict(sd: Dict[str, torch.Tensor], g_gain: float = 2.2, refiner_gain: float = 2.35,
                       proj_gain: float = 8.0, head_gain: float = 3.0, mel_offset: float = -4.0) -> Dict[str, torch.Tensor]:
    """A copy of ``sd`` with larger weights, for parity fixtures at realistic magnitudes.

    Random-init keeps every activation below ~0.5, where APTx is nearly linear and errors are tiny.  Here every
    weight-norm gain ``g`` (both flavours) is multiplied by ``g_gain`` (``refiner_gain`` inside the refiner), ``proj``
    by ``proj_gain`` (ConvBlock2D's stencil value then leaves the +-64 range of its table), ``q_out_proj`` /
    ``out_proj`` / ``hidden_proj`` by ``head_gain`` and ``out_proj.bias`` is shifted by ``mel_offset``: activations reach
    |x| ~ 5-60 (tanh saturated) and the re-encoded mels the log-mel range of convert_spectrograms.py:34."""
    out = {}
    for k, v in sd.items():
        t = v.clone()
        if k.endswith("original0") or k.endswith("weight_g"):
            t = t * (refiner_gain if k.startswith("refiner.") else g_gain)
        elif k in ("proj.weight", "proj.bias"):
            t = t * proj_gain
        elif k in ("q_out_proj.weight", "out_proj.weight", "hidden_proj.weight"):
            t = t * head_gain
        elif k == "out_proj.bias":
            t = t + mel_offset
        out[k] = t
    return out
