"""Host logic of the training driver (mqgan_b200/train_cli.py): the reference's dataset / collate / split /
checkpoint-naming behaviour (train.py:84-198, 243-257, 339-343, 660-679).  CPU tests, plus one GPU smoke run."""
import json
import os
import random

import numpy as np
import pytest
import torch
import yaml

from mqgan_b200 import train_cli as TC


def _write_mels(root, n, n_mels, lens):
    os.makedirs(os.path.join(root, "sub"), exist_ok=True)
    g = np.random.default_rng(0)
    for i in range(n):
        d = root if i % 2 else os.path.join(root, "sub")
        np.save(os.path.join(d, f"utt{i:03d}.npy"), (g.standard_normal((lens[i % len(lens)], n_mels)) * 2 - 4).astype(np.float32))


def test_dataset_walks_recursively_and_crops(tmp_path):
    _write_mels(str(tmp_path), 6, 8, [30, 50, 20])
    ds = TC.RealMelSpectrogramDataset(str(tmp_path), crop_len=None)
    assert len(ds) == 6 and all(f.endswith(".npy") for f in ds.filenames)
    mel, n, name = ds[0]
    assert mel.dtype == torch.float32 and mel.shape == (n, 8) and name.startswith("utt")
    ds40 = TC.RealMelSpectrogramDataset(str(tmp_path), crop_len=40)
    for i in range(6):
        mel, n, _ = ds40[i]
        assert mel.shape == (40, 8) and n == min(40, ds[i][1])
        if n < 40:
            assert float(mel[n:].abs().max()) == 0.0
    (tmp_path / "broken.npy").write_bytes(b"not a numpy file")
    ds2 = TC.RealMelSpectrogramDataset(str(tmp_path), None)
    assert any(ds2[i] is None for i in range(len(ds2)))             # unreadable files are skipped by the collate fn


def test_collate_one_length_per_batch_pad_and_crop():
    random.seed(0)
    items = [(torch.ones(50, 4), 50, "a"), (torch.ones(20, 4) * 2, 20, "b"), None, (torch.ones(32, 4) * 3, 32, "c")]
    x, lens, names = TC.pad_collate_fn(items, crop_lens=32)
    assert x.shape == (3, 32, 4) and lens.tolist() == [32, 20, 32] and lens.dtype == torch.int32 and names == ("a", "b", "c")
    assert float(x[1, 20:].abs().max()) == 0.0 and float(x[1, :20].min()) == 2.0
    seen = set()
    for _ in range(20):
        x, lens, _ = TC.pad_collate_fn(items, crop_lens=[16, 24, 64])
        seen.add(x.shape[1])
        assert lens.max() <= x.shape[1]
    assert seen == {16, 24, 64}
    x, lens, _ = TC.pad_collate_fn(items, crop_lens=None)           # pad to the batch maximum
    assert x.shape == (3, 50, 4) and lens.tolist() == [50, 20, 32]
    assert TC.pad_collate_fn([None, None]) == (None, None, None)


def test_split_is_seeded_and_sized():
    tr, ev = TC.split_dataset(100, 0.02, seed=42)
    assert len(tr) == 98 and len(ev) == 2 and sorted(tr + ev) == list(range(100))
    assert (tr, ev) == TC.split_dataset(100, 0.02, seed=42)
    assert tr != TC.split_dataset(100, 0.02, seed=43)[0]
    with pytest.raises(ValueError):
        TC.split_dataset(1, 1.0, seed=0)


def test_cli_flags_are_the_references():
    p = TC.build_arg_parser()
    a = p.parse_args([])
    assert (a.config, a.pretrained, a.output_dir) == ("config.yaml", None, None)       # train.py:662-664
    a = p.parse_args(["--config", "c.yaml", "--pretrained", "p.pth", "--output_dir", "o"])
    assert (a.config, a.pretrained, a.output_dir) == ("c.yaml", "p.pth", "o")


def _tiny_config(data_dir, out_dir):
    return {
        "project_name": "MQGAN",
        "data": {"data_dir": data_dir, "output_dir": out_dir, "validation_split": 0.1, "crop_len": [32, 48], "batch_size": 4,
                 "num_workers": 0},
        "model": {"mel_channels": 32,
                  "generator": {"channels": [64, 64, 64, 128], "kernel_sizes": [3, 3, 5, 7], "dropout": 0.1, "fsq_levels": [8, 5, 5, 5],
                                "refiner_base_channels": 16, "refiner_depth": 3},
                  "discriminator_patch": {"hidden_channels": [16, 16, 32], "kernel_sizes": [5, 5, 3, 3],
                                          "strides": [[1, 2], [2, 2], [2, 1], [1, 1]]},
                  "discriminator_multibin": {"hidden_channels": [16, 16, 32], "kernel_sizes": [7, 5, 3, 3], "n_bins": 2,
                                             "n_no_strides": 2}},
        "training": {"num_epochs": 3, "lr": 1e-4, "beta1": 0.9, "beta2": 0.999, "lr_d_factor": 1.15, "d_beta1": 0.5, "d_beta2": 0.999,
                     "warmup_steps": 10, "discriminator_train_start_epoch": 2,
                     "loss_weights": {"fm_lambda": 0.25, "Gloss_lambda": 15.0, "recon_lambda": 15.0}, "use_fm_loss": False, "seed": 42,
                     "no_cuda": False, "pretrained": None},
        "logging": {"eval_interval": 2, "save_interval": 1, "num_plot_examples": 2, "wandb": {"entity": None, "project": "MQGAN"}},
    }


@pytest.mark.gpu
def test_train_cli_runs_checkpoints_and_resumes(tmp_path):
    """Three tiny epochs through the drop-in ``train.py`` entry point (reconstruction-only epoch, then GAN epochs with
    CUDA-graph replay), the checkpoint loads into the inference PreEncoder, and a second run resumes after it."""
    data, out = str(tmp_path / "mels"), str(tmp_path / "run")
    _write_mels(data, 18, 32, [40, 60, 25, 90])
    cfg_path = str(tmp_path / "cfg.yaml")
    with open(cfg_path, "w") as f:
        yaml.safe_dump(_tiny_config(data, out), f)
    assert TC.main(["--config", cfg_path]) == 0
    ckpts = sorted(os.listdir(out))
    assert [c for c in ckpts if c.endswith(".pth")] == ["checkpoint_epoch_001.pth", "checkpoint_epoch_002.pth", "checkpoint_epoch_003.pth"]
    log = [json.loads(l) for l in open(os.path.join(out, "train_log.jsonl"))]
    assert len(log) == 15 and log[0]["loss_d"] == 0.0 and log[-1]["loss_d"] > 0.0             # 17 training files = 5 batches x 3 epochs; GAN from epoch 2
    assert all(np.isfinite(list(r.values())).all() for r in log)
    assert log[1]["learning_rate"] > log[0]["learning_rate"]                                    # warm-up
    ck = torch.load(os.path.join(out, "checkpoint_epoch_003.pth"), map_location="cpu", weights_only=False)
    assert {"epoch", "model_state_dict", "optimizer_g_state_dict", "optimizer_d_state_dict", "scaler_g_state_dict",
            "scaler_d_state_dict", "config"} <= set(ck)
    from mqgan_b200.preencoder import get_pre_encoder
    m = get_pre_encoder(os.path.join(out, "checkpoint_epoch_003.pth"), "cuda", channels=[64, 64, 64, 128], kernel_sizes=[3, 3, 5, 7],
                        mel_channels=32, fsq_levels=[8, 5, 5, 5], refiner_base_channels=16, refiner_depth=3, inference=True)
    idx = m.encode(torch.randn(2, 40, 32, device="cuda") * 2 - 4)
    assert idx.shape == (2, 40) and idx.dtype == torch.int64
    # resume: a fourth epoch only
    cfg = _tiny_config(data, out)
    cfg["training"]["num_epochs"] = 4
    with open(cfg_path, "w") as f:
        yaml.safe_dump(cfg, f)
    assert TC.main(["--config", cfg_path]) == 0
    assert os.path.exists(os.path.join(out, "checkpoint_epoch_004.pth"))
    log2 = [json.loads(l) for l in open(os.path.join(out, "train_log.jsonl"))]
    assert len(log2) == 20 and log2[-1]["epoch"] == 4
    # a resumed trainer's optimisers read the LIVE learning-rate tensors (load_state_dict replaces them with copies):
    # the warm-up schedule must keep reaching Adam after a resume
    t = TC.Trainer(cfg, max_steps=1)
    assert t.start_epoch == 5
    assert all(g["lr"] is t.step.lr_g for g in t.step.opt_g.param_groups)
    assert all(g["lr"] is t.step.lr_d for g in t.step.opt_d.param_groups)
    t.step.lr_g.fill_(0.125)
    assert float(t.step.opt_g.param_groups[0]["lr"]) == 0.125


def test_epoch_schedule_is_lockstep_across_replicas():
    idx = list(range(103))
    world = 4
    per_rank = [TC.epoch_schedule(idx, 8, seed=42, epoch=3, rank=r, world=world, crop_len=[128, 192, 256]) for r in range(world)]
    n_it = {len(s) for s in per_rank}
    assert n_it == {3}                                              # 13 batches -> 12 usable -> 3 iterations each
    for it in range(3):
        assert len({per_rank[r][it][1] for r in range(world)}) == 1            # same crop length on every replica
    seen = [j for s in per_rank for b, _ in s for j in b]
    assert len(seen) == len(set(seen)) == 96                        # disjoint batches
    assert per_rank[1] == TC.epoch_schedule(idx, 8, 42, 3, 1, world, [128, 192, 256])          # deterministic
    assert per_rank[1] != TC.epoch_schedule(idx, 8, 42, 4, 1, world, [128, 192, 256])          # reshuffled every epoch
    single = TC.epoch_schedule(idx, 8, 42, 3, 0, 1, 256)
    assert len(single) == 13 and sorted(j for b, _ in single for j in b) == idx and {t for _, t in single} == {256}
    assert TC.epoch_schedule(idx, 8, 42, 3, 0, 1, None)[0][1] is None
