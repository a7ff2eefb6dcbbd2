"""Runs the UNMODIFIED reference (ZDisket/MQGAN, pure Python on PyTorch) as the CPU arm.  TEST / BENCH
INFRASTRUCTURE ONLY: imported by ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, ``__graft_entry__.build``
and the golden generators - never by anything under ``mqgan_b200/``.

The reference is a directory of bare top-level modules.  ``install()`` copies its ``*.py`` files and
``configs/`` from ``/root/reference`` into ``baseline/_ref/`` (git-ignored: the install is not product source and never
enters history; it is NOT gpurun-ignored, so it travels to the GPU box like the built ``.so``), SURVEY §8(c).  The one
missing dependency of the hot path, ``einx`` (used only in FSQ's training branch, quantizer.py:151,160), is stubbed before
import; nothing else is touched.
"""
from __future__ import annotations

import glob
import os
import shutil
import sys
import types
from typing import Optional

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALL_DIR = os.path.join(ROOT, "baseline", "_ref")
SOURCE_DIR = "/root/reference"


def install(src: str = SOURCE_DIR, dst: str = INSTALL_DIR) -> Optional[str]:
    """Copy the reference's Python modules + configs into baseline/_ref (no-op without ``src``)."""
    if not os.path.isfile(os.path.join(src, "preencoder.py")):
        return None
    os.makedirs(dst, exist_ok=True)
    for f in glob.glob(os.path.join(src, "*.py")) + glob.glob(os.path.join(src, "requirements.txt")):
        shutil.copy2(f, os.path.join(dst, os.path.basename(f)))
    if os.path.isdir(os.path.join(src, "configs")):
        shutil.copytree(os.path.join(src, "configs"), os.path.join(dst, "configs"), dirs_exist_ok=True)
    return dst


def find() -> Optional[str]:
    """Directory holding the reference's modules: $MQGAN_REFERENCE, baseline/_ref, or /root/reference."""
    for d in (os.environ.get("MQGAN_REFERENCE"), INSTALL_DIR, SOURCE_DIR):
        if d and os.path.isfile(os.path.join(d, "preencoder.py")):
            return d
    return None


def import_preencoder():
    """The reference's ``preencoder`` module (with the einx stub), or None when no copy is reachable."""
    d = find()
    if d is None:
        return None
    einx = types.ModuleType("einx")
    einx.where = lambda pattern, cond, a, b: torch.where(cond.view(-1, *([1] * (a.dim() - 1))), a, b)
    sys.modules.setdefault("einx", einx)
    if d not in sys.path:
        sys.path.insert(0, d)
    import preencoder as ref_pre  # noqa: the reference's module, not mqgan_b200.preencoder
    if os.path.dirname(os.path.abspath(ref_pre.__file__)) != os.path.abspath(d):
        raise ImportError(f"'preencoder' resolved to {ref_pre.__file__}, not the reference in {d}")
    return ref_pre


def build_model(ref_pre, cfg, sd):
    """reference PreEncoder(cfg) in eval mode with the given state-dict (preencoder.py:305-361)."""
    m = ref_pre.PreEncoder(cfg.mel_channels, list(cfg.channels), list(cfg.kernel_sizes),
                           fsq_levels=list(cfg.fsq_levels), dropout=0.0,
                           refiner_base_channels=cfg.refiner_base_channels, refiner_depth=cfg.refiner_depth,
                           refiner_hidden_proj_divisor=cfg.refiner_hidden_proj_divisor)
    m.load_state_dict(sd, strict=True)
    return m.eval()


def reencode(ref_pre, model, mel: torch.Tensor, lengths: torch.Tensor, micro_batch: Optional[int] = None):
    """encode + decode as reencode_spectrograms_from_checkpoint.py:76-86 drives them (mask from lengths), optionally
    in micro-batches of utterances (full-length utterances are batch-invariant)."""
    B = mel.shape[0]
    mb = B if not micro_batch else max(1, int(micro_batch))
    idxs, outs = [], []
    with torch.no_grad():
        for b0 in range(0, B, mb):
            x, l = mel[b0:b0 + mb], lengths[b0:b0 + mb]
            mask = ref_pre.sequence_mask(x.shape[1], l).unsqueeze(1)
            idx = model.encode(x, mask)
            idxs.append(idx)
            outs.append(model.decode(idx, mask))
    return torch.cat(idxs), torch.cat(outs)
