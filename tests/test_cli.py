"""The drop-in CLIs end to end: a TorchScript model directory / a .pth checkpoint, a tree of
ragged .npy mels, and the reference's batching.  GPU tests compare every written file with the
oracle run on the same batches (batch composition matters: SURVEY App. B3)."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import yaml

from mqgan_b200 import spec as S, reencode as R
from mqgan_b200.engine import folded_weights
from mqgan_b200.preencoder import _Node, _attach
from mqgan_b200.scripted_preencoder import ScriptedPreEncoder
from oracle import preencoder_oracle as O
from tests.helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _yaml_cfg(cfg):
    return {"model": {"mel_channels": cfg.mel_channels, "generator": {
        "channels": list(cfg.channels), "kernel_sizes": list(cfg.kernel_sizes), "dropout": 0.0,
        "fsq_levels": list(cfg.fsq_levels), "refiner_base_channels": cfg.refiner_base_channels,
        "refiner_depth": cfg.refiner_depth}}}


def _export_like_reference(sd, model_dir, cfg):
    """What convert_to_torchscript.py leaves on disk, as far as this loader can see: a TorchScript
    module whose state_dict has the parametrised keys plus plain decoder conv weights (legacy
    weight-norm stripped by inference=True), and model_config.yaml."""
    w = folded_weights(sd)
    root = _Node()
    for k, v in sd.items():
        if k.endswith("weight_g"):
            continue
        if k.endswith("weight_v"):
            _attach(root, k[:-2], nn.Parameter(w[k[:-2]].clone()))
        else:
            _attach(root, k, nn.Parameter(v.clone()))
    os.makedirs(model_dir, exist_ok=True)
    torch.jit.script(root).save(os.path.join(model_dir, "model_cpu.pt"))
    with open(os.path.join(model_dir, "model_config.yaml"), "w") as f:
        yaml.safe_dump(_yaml_cfg(cfg), f)


def _tree(root, n_mels, n=7, seed=0):
    rng = np.random.default_rng(seed)
    for i in range(n):
        d = os.path.join(root, f"spk{i % 2}")
        os.makedirs(d, exist_ok=True)
        T = int(rng.integers(20, 70))
        np.save(os.path.join(d, f"u{i}.npy"), (rng.standard_normal((T, n_mels)) * 2 - 4).astype(np.float32))


def test_scripted_preencoder_error_contract(tmp_path):
    with pytest.raises(FileNotFoundError):
        ScriptedPreEncoder(str(tmp_path / "nope"), device="cuda")
    d = tmp_path / "m"
    d.mkdir()
    with pytest.raises(FileNotFoundError, match="model_config.yaml"):
        ScriptedPreEncoder(str(d), device="cuda")
    (d / "model_config.yaml").write_text(yaml.safe_dump(_yaml_cfg(S.TINY)))
    with pytest.raises(RuntimeError, match="CUDA"):
        ScriptedPreEncoder(str(d), device="cpu")           # no CPU fallback in this build
    if torch.cuda.is_available():
        with pytest.raises(FileNotFoundError):
            ScriptedPreEncoder(str(d), device="cuda")      # no model_cuda.pt / model_cpu.pt


def _check_outputs(cfg, sd, src, dst, batch_size):
    files = R.list_npy_files(src)
    w = O.effective_weights(sd)
    worst = 0.0
    mism = 0
    for batch in R.make_batches(files, batch_size):
        mel, lengths = R.load_and_pad(batch)
        mask = O.sequence_mask(mel.shape[1], torch.tensor(lengths)).unsqueeze(1)
        idx = O.encode(w, cfg, mel, mask, folded=True)
        ref = O.decode(w, cfg, idx, mask, folded=True)
        for i, p in enumerate(batch):
            out = np.load(os.path.join(dst, os.path.relpath(p, src)))
            assert out.dtype == np.float32 and out.shape == (lengths[i], cfg.mel_channels)
            worst = max(worst, float(np.abs(out - ref[i, : lengths[i]].numpy()).max()))
    return worst


@pytest.mark.gpu
def test_reencode_cli_torchscript_dir(tmp_path):
    import reencode_spectrograms as cli
    cfg, sd, _, _, _ = load_golden("tiny")
    model_dir, src, dst = str(tmp_path / "model"), str(tmp_path / "in"), str(tmp_path / "out")
    _export_like_reference(sd, model_dir, cfg)
    _tree(src, cfg.mel_channels)
    cli.reencode_spectrograms(model_dir, src, dst, "cuda", 3)
    assert len(R.list_npy_files(dst)) == 7
    worst = _check_outputs(cfg, sd, src, dst, 3)
    assert worst < 2e-2, worst           # same mel tolerance as tests/test_gpu_parity.py
    # ScriptedPreEncoder API details
    m = ScriptedPreEncoder(model_dir, device="cuda")
    assert m.mel_channels == cfg.mel_channels and m.fsq_levels == list(cfg.fsq_levels)
    with pytest.raises(ValueError):
        m.encode(torch.zeros(4, cfg.mel_channels))
    x = torch.randn(2, 33, cfg.mel_channels) * 2 - 4
    idx = m.encode(x, lengths=[33, 20])
    assert idx.shape == (2, 33) and idx.dtype == torch.int64
    assert m.decode(idx, lengths=torch.tensor([33, 20])).shape == (2, 33, cfg.mel_channels)
    assert m.decode(m.encode(x)).shape == x.shape          # lengths=None works here (fails in the reference, B9)


@pytest.mark.gpu
def test_reencode_cli_from_checkpoint(tmp_path):
    import reencode_spectrograms_from_checkpoint as cli
    cfg, sd, _, _, _ = load_golden("tiny")
    src, dst = str(tmp_path / "in"), str(tmp_path / "out")
    ck = str(tmp_path / "ck.pth")
    torch.save({"model_state_dict": {"module." + k: v for k, v in sd.items()}, "epoch": 3}, ck)
    _tree(src, cfg.mel_channels, n=5, seed=2)
    cli.reencode_spectrograms(ck, _yaml_cfg(cfg), src, dst, "cuda", 2)
    assert _check_outputs(cfg, sd, src, dst, 2) < 2e-2
