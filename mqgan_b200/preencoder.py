"""Drop-in ``PreEncoder`` boundary (reference: preencoder.py:304-599).

Same constructor, attributes, state-dict keys (both weight-norm flavours,
SURVEY App. B4), ``encode`` / ``decode`` / ``forward`` signatures and loader as
the reference, but the arithmetic runs in libmqgan_b200.so on a B200.  The
module owns ordinary ``nn.Parameter``s so ``load_state_dict(strict=True)`` /
``state_dict()`` / ``.to()`` behave as users expect; kernels read a packed copy
that is rebuilt whenever a parameter changes.

There is no CPU path: calling ``encode`` / ``decode`` on a module that is not on
a CUDA device raises.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from typing import List, Optional, Sequence

import threading

import torch
import torch.nn as nn

from .engine import PreEncoderEngine
from .spec import PreEncoderConfig, param_spec


_ENGINE_LOCK = threading.Lock()


def sequence_mask(max_length, x_lengths):
    """(B, max_length) bool, True = padded (preencoder.py:15-24)."""
    ar = torch.arange(max_length, device=x_lengths.device)
    return ar.unsqueeze(0) >= x_lengths.unsqueeze(1)


class _Node(nn.Module):
    """Bare container so dotted reference key names resolve to real parameters."""


def _attach(root: nn.Module, dotted: str, param: nn.Parameter) -> None:
    parts = dotted.split(".")
    mod = root
    for name in parts[:-1]:
        nxt = mod._modules.get(name)
        if nxt is None:
            nxt = _Node()
            mod.add_module(name, nxt)
        mod = nxt
    mod.register_parameter(parts[-1], param)


class PreEncoder(nn.Module):
    def __init__(self, mel_channels, channels, kernel_sizes, fsq_levels=[8, 8, 5, 5, 5], dropout=0.1,
                 refiner_base_channels=128, refiner_depth=3, refiner_hidden_proj_divisor=8,
                 encoder_precision: str = "f16x2", decoder_precision: str = "bf16"):
        super().__init__()
        self.cfg = PreEncoderConfig(int(mel_channels), tuple(channels), tuple(kernel_sizes), tuple(fsq_levels),
                                    int(refiner_base_channels), int(refiner_depth),
                                    int(refiner_hidden_proj_divisor))
        self.dropout_p = dropout            # eval-only path: dropout is the identity
        self.encoder_precision = encoder_precision
        # additive: "bf16" (default, fastest) or "f16x2" = fp32-grade decoder / refiner (the reference's decode is fp32)
        self.decoder_precision = decoder_precision
        # attributes the reference exposes (preencoder.py:323, 337-341, 355)
        self.quantizer_dim = self.cfg.quantizer_dim
        self.codebook_size = self.cfg.codebook_size
        self.bos_token_id = self.codebook_size + 1
        self.eos_token_id = self.codebook_size + 2
        self.refiner_hidden_channels = self.cfg.refiner_hidden_channels
        spec = param_spec(self.cfg)
        shapes = dict(spec)
        for key, shape in spec:
            _attach(self, key, nn.Parameter(self._init(key, shape, shapes)))
        # weight-norm g starts as ||v|| (what weight_norm does when it wraps a conv)
        with torch.no_grad():
            sd = dict(self.named_parameters())
            for key, _ in spec:
                if key.endswith("original0"):
                    v = sd[key[: -len("original0")] + "original1"]
                elif key.endswith("weight_g"):
                    v = sd[key[: -len("weight_g")] + "weight_v"]
                else:
                    continue
                sd[key].copy_(v.reshape(v.shape[0], -1).norm(dim=1).reshape(sd[key].shape))
        self._engine: Optional[PreEncoderEngine] = None
        self._engine_key = None
        self.eval()

    @staticmethod
    def _init(key, shape, shapes):
        if key.endswith(".relu.beta"):
            return torch.tensor(1.0)
        if key.endswith(".relu.gamma"):
            return torch.tensor(0.5)
        wshape = shape
        if key.endswith(".bias"):
            base = key[: -len(".bias")]
            for cand in (base + ".weight", base + ".parametrizations.weight.original1", base + ".weight_v"):
                if cand in shapes:
                    wshape = shapes[cand]
                    break
        fan_in = 1
        for d in wshape[1:]:
            fan_in *= d
        bound = 1.0 / math.sqrt(max(fan_in, 1))
        return (torch.rand(shape) * 2.0 - 1.0) * bound

    # ------------------------------------------------------------------
    def _device(self) -> torch.device:
        return next(self.parameters()).device

    def engine(self) -> PreEncoderEngine:
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("mqgan_b200.PreEncoder runs on CUDA (B200) only - there is no CPU fallback; "
                               "move the module with .to('cuda')")
        key = (dev, self.encoder_precision, self.decoder_precision) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._engine is None or key != self._engine_key:
            with _ENGINE_LOCK:                  # the CLI drives one model from two compute threads
                if self._engine is None or key != self._engine_key:
                    self._engine = PreEncoderEngine(self.cfg, self.state_dict(), dev, self.encoder_precision,
                                                    decoder_precision=self.decoder_precision)
                    self._engine_key = key
        return self._engine

    @torch.no_grad()
    def encode(self, x, x_mask=None):
        """(B, T, mel) [+ (B,1,T) bool, True = padded] -> (B, T) int64 (preencoder.py:420-451)."""
        return self.engine().encode(x.to(self._device()), x_mask)

    @torch.no_grad()
    def decode(self, indices, x_mask=None, return_hidden=False, *, host_out=None, lengths=None):
        """(B, T) int -> (B, T, mel) [, (B, C0, T)] (preencoder.py:453-504).  Additive, keyword-only: ``host_out``, a
        pinned host tensor that receives the result chunk by chunk while later chunks are still computing;
        ``lengths``, the utterance lengths as HOST integers matching ``x_mask`` - a ragged batch then goes through the
        refiner in length-sorted groups (identical output, less padding computed)."""
        if lengths is not None and x_mask is None:
            raise ValueError("decode(lengths=...) describes the padding of x_mask; pass both")
        if lengths is not None:
            lengths = [int(v) for v in (lengths.tolist() if isinstance(lengths, torch.Tensor) else lengths)]
        return self.engine().decode(indices.to(self._device()), x_mask, return_hidden=return_hidden, host_out=host_out,
                                    lengths_host=lengths)

    def forward(self, x, x_lengths):
        """(x_recon, x_post) as preencoder.py:363-418.

        In eval mode / under ``torch.no_grad()`` this is encode + decode on the inference engine.  In training mode
        with autograd on it is the differentiable training forward of ``mqgan_b200.training`` (tcgen05 forward /
        data-gradient / weight-gradient convolutions, fused ConvBlock2D and activation passes, straight-through FSQ)
        over this module's own parameters, so ``loss.backward()`` fills their ``.grad`` and any optimiser can step
        them - what the reference's ``Trainer`` does with ``self.generator(real, lens)`` (train.py:524).  Dropout is not
        applied (see training.py); a non-zero ``dropout`` argument is reported once and ignored."""
        if self.training and torch.is_grad_enabled():
            from . import training as _training
            if self.dropout_p and not getattr(self, "_dropout_warned", False):
                print(f"Warning: PreEncoder(dropout={self.dropout_p}) - the B200 training forward runs without dropout.")
                self._dropout_warned = True
            dev = self._device()
            if dev.type != "cuda":
                raise RuntimeError("mqgan_b200.PreEncoder runs on CUDA (B200) only - there is no CPU fallback")
            params = dict(self.named_parameters())
            with torch.cuda.device(dev):               # launches go to the current device's stream
                return _training.generator_forward(params, self.cfg, x.to(dev), x_lengths.to(dev))
        with torch.no_grad(), torch.cuda.device(self._device()):
            x = x.to(self._device())
            mask = sequence_mask(x.size(1), x_lengths.to(x.device)).unsqueeze(1)
            eng = self.engine()
            idx = eng.encode(x, mask)
            x_post, x_recon = eng.decode(idx, mask, return_recon=True)
            return x_recon, x_post


def strip_weight_norm(module):
    """Reference: preencoder.py:507-514.  Weight-norm is folded once when the
    kernels' packed weights are built, so there is nothing to strip; kept for
    call-site compatibility."""
    return module


def get_pre_encoder(model_path: str, device, channels=[384, 512, 768], kernel_sizes=[7, 5, 3],
                    mel_channels=88, fsq_levels=[8, 5, 5, 5], refiner_base_channels=128, refiner_depth=3,
                    refiner_hidden_proj_divisor=8, inference=False):
    """Load a checkpoint written by the reference's train.py (preencoder.py:517-599):
    needs ``checkpoint['model_state_dict']``, strips a ``module.`` prefix, strict load,
    ``.to(device).eval()``.  Raises FileNotFoundError / KeyError / RuntimeError as the reference."""
    if not os.path.isfile(model_path):
        raise FileNotFoundError(f"Checkpoint file not found: {model_path}")
    print(f"Loading checkpoint from: {model_path}")
    checkpoint = torch.load(model_path, map_location="cpu", weights_only=False)
    try:
        model = PreEncoder(mel_channels=mel_channels, channels=channels, kernel_sizes=kernel_sizes, dropout=0.0,
                           fsq_levels=fsq_levels, refiner_base_channels=refiner_base_channels,
                           refiner_depth=refiner_depth, refiner_hidden_proj_divisor=refiner_hidden_proj_divisor)
    except Exception as e:
        raise RuntimeError(f"Failed to instantiate model with loaded config: {e}")
    if "model_state_dict" not in checkpoint:
        raise KeyError("Checkpoint missing 'model_state_dict' key containing weights.")
    weights = OrderedDict()
    stripped = False
    for k, v in checkpoint["model_state_dict"].items():
        if k.startswith("module."):
            stripped = True
            k = k[7:]
        weights[k] = v
    if stripped:
        print("Removed 'module.' prefix from weight keys.")
    weights = _accept_stripped_weight_norm(model, weights)
    try:
        model.load_state_dict(weights, strict=True)
        print("Successfully loaded model weights.")
    except RuntimeError as e:
        print(f"Error loading state_dict (likely architecture mismatch): {e}")
        raise
    model.to(device)
    model.eval()
    if inference:
        strip_weight_norm(model)
    print(f"Model loaded onto {device} and set to evaluation mode.")
    return model


def _accept_stripped_weight_norm(model: "PreEncoder", weights):
    """A state-dict harvested from an ``inference=True`` reference model / TorchScript
    export has plain ``decoder_blocks.*.conv*.weight`` (legacy weight-norm removed,
    SURVEY App. B4).  Re-express it as g = ||w||, v = w so the strict load still holds."""
    own = set(k for k, _ in model.named_parameters())
    out = OrderedDict()
    for k, v in weights.items():
        if k not in own and k.endswith(".weight") and (k[: -len("weight")] + "weight_v") in own:
            base = k[: -len("weight")]
            out[base + "weight_v"] = v
            out[base + "weight_g"] = v.reshape(v.shape[0], -1).norm(dim=1).reshape(v.shape[0], *([1] * (v.dim() - 1)))
        else:
            out[k] = v
    return out
