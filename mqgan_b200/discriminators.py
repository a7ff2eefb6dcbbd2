"""Drop-in ``MelSpectrogramPatchDiscriminator2D`` / ``MultiBinDiscriminator`` (reference: discriminators.py:70-312).

Same constructor arguments, ``forward`` signatures / return values and state-dict keys (legacy ``spectral_norm``:
``convs.N.weight_orig`` / ``bias`` parameters, ``weight_u`` / ``weight_v`` buffers; ``se_block.fc1`` / ``fc2``) as the
reference, so a training script only swaps the import.  The arithmetic is ``mqgan_b200.training.patch_discriminator``:
on a CUDA device with ``fast=True`` (default there) the feature maps stay bf16 / channels_last as under the reference's
autocast (train.py:523) and the bias + LeakyReLU + patch-mask passes run in the library; otherwise plain fp32.
"""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import training as _training
from .spec import PatchDiscConfig, is_disc_buffer, patch_disc_param_spec


def _layer_kernels(kernel_sizes, lengthwise_only: bool) -> Tuple[Tuple[int, int], ...]:
    out = []
    for k in kernel_sizes:
        k1, k2 = (k if isinstance(k, (tuple, list)) else (k, k))
        out.append((1, int(k2)) if lengthwise_only else (int(k1), int(k2)))          # discriminators.py:128-143, 160-170
    return tuple(out)


def _layer_strides(stride, n: int, lengthwise_only: bool) -> Tuple[Tuple[int, int], ...]:
    if isinstance(stride, int):                                                      # :115-116
        strides = [(1, stride)] * n
    elif isinstance(stride, tuple) and len(stride) == 2 and not isinstance(stride[0], (tuple, list)):
        strides = [tuple(stride)] * n                                                # :117-118
    else:
        if len(stride) != n:
            raise AssertionError("stride list must match kernel_sizes")              # :121
        strides = [tuple(s) for s in stride]
    if lengthwise_only:
        strides = [(1, s[1]) for s in strides]
    return tuple((int(a), int(b)) for a, b in strides)


class _Holder(nn.Module):
    """Bare container so the reference's dotted key names resolve to real parameters / buffers."""


class MelSpectrogramPatchDiscriminator2D(nn.Module):
    def __init__(self, mel_channels: int, hidden_channels: list = (64, 128, 256, 512), kernel_sizes: list = (7, 5, 5, 3, 3),
                 stride: Union[int, Tuple[int, int], List[Tuple[int, int]]] = (2, 2), lengthwise_only=False, *,
                 fast: Optional[bool] = None):
        super().__init__()
        assert len(kernel_sizes) == len(hidden_channels) + 1, "kernel_sizes must be hidden_channels len + 1"     # :99-101
        self.mel_channels = mel_channels
        self.fast = fast
        self.cfg = PatchDiscConfig(int(mel_channels), tuple(int(h) for h in hidden_channels),
                                   _layer_kernels(kernel_sizes, lengthwise_only),
                                   _layer_strides(stride, len(kernel_sizes), lengthwise_only))
        self.ret_features_map = list(self.cfg.feature_layers)
        self.convs = nn.ModuleList(_Holder() for _ in self.cfg.kernels)
        self.se_block = _Holder()
        self.se_block.fc1, self.se_block.fc2 = _Holder(), _Holder()
        for key, shape in patch_disc_param_spec(self.cfg):
            mod = self
            parts = key.split(".")
            for name in parts[:-1]:
                mod = mod[int(name)] if name.isdigit() else getattr(mod, name)
            if key.endswith(".weight_orig"):
                t = torch.empty(shape).normal_(0.0, 0.02)                             # _initialize_weights :193-198
            elif is_disc_buffer(key):
                t = F.normalize(torch.randn(shape), dim=0, eps=1e-12)                 # spectral_norm's u / v
            elif key.endswith(".bias") and ".convs." in "." + key:
                t = torch.zeros(shape)
            else:                                                                     # squeeze-excite linears: nn.Linear default
                fan_in = shape[1] if len(shape) == 2 else dict(patch_disc_param_spec(self.cfg))[key[:-4] + "weight"][1]
                t = (torch.rand(shape) * 2 - 1) / fan_in ** 0.5
            if is_disc_buffer(key):
                mod.register_buffer(parts[-1], t)
            else:
                mod.register_parameter(parts[-1], nn.Parameter(t))

    def _state(self):
        sd = dict(self.named_parameters())
        sd.update(dict(self.named_buffers()))
        return sd

    def forward(self, x: torch.Tensor, x_lengths: torch.Tensor, return_features: bool = False):
        """x (B, T, F), x_lengths (B,) -> (logits (B,1,H,W), patch_mask True = valid[, features]) (:208-257)."""
        fast = x.is_cuda if self.fast is None else (self.fast and x.is_cuda)
        out, patch_mask, feats = _training.patch_discriminator(self._state(), self.cfg, x, x_lengths, self.training,
                                                               autocast_bf16=fast)
        if return_features:
            return out, patch_mask, feats
        return out, patch_mask


class MultiBinDiscriminator(nn.Module):
    """discriminators.py:260-312: ``n_bins`` equal mel bands, an independent patch discriminator on each."""

    def __init__(self, mel_channels: int, n_bins: int = 4, hidden_channels: list = (64, 128, 256, 512),
                 kernel_sizes: list = (7, 5, 5, 3, 3), n_no_strides: int = 2, *, fast: Optional[bool] = None):
        super().__init__()
        assert mel_channels % n_bins == 0, "mel_channels must divide n_bins"
        for h in hidden_channels:
            assert h % n_bins == 0, f"hidden size {h} must divide n_bins"
        self.n_bins = n_bins
        strides = [(1, 1) if i < n_no_strides else (1, 2) for i in range(len(kernel_sizes))]
        self.discriminators = nn.ModuleList(
            MelSpectrogramPatchDiscriminator2D(mel_channels // n_bins, list(hidden_channels), [(3, ks) for ks in kernel_sizes],
                                               stride=strides, fast=fast)
            for _ in range(n_bins))

    def forward(self, x: torch.Tensor, x_lengths: torch.Tensor, return_features: bool = False):
        splits = torch.split(x, x.size(-1) // self.n_bins, dim=-1)
        if x.is_cuda and self.n_bins > 1:
            # the bands are independent and individually too small for the GPU: one CUDA stream each (forked from and
            # joined to the caller's stream; autograd replays every band's backward on its stream)
            cur = torch.cuda.current_stream(x.device)
            streams = _training._side_streams(x.device, self.n_bins)
            res = []
            for disc, sub, st in zip(self.discriminators, splits, streams):
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    res.append(disc(sub, x_lengths, True))
            for st in streams:
                cur.wait_stream(st)
        else:
            res = [disc(sub, x_lengths, True) for disc, sub in zip(self.discriminators, splits)]
        outs, masks, feats = [r[0] for r in res], [r[1] for r in res], [r[2] for r in res]
        if return_features:
            return outs, masks, feats
        return outs, masks
