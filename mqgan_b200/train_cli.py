"""Drop-in for the reference's ``train.py`` driver around the training step: same command line
(``--config --pretrained --output_dir``, train.py:660-679), same YAML schema (configs/model_config_*.yaml), same
dataset behaviour (recursive ``*.npy`` mels, seeded validation split, ONE crop length drawn per batch, random
crop / zero pad: train.py:84-198, 243-270), same checkpoint files (``checkpoint_epoch_NNN.pth`` with
``model_state_dict`` in the reference's key names - loadable by the reference's and this repo's
``get_pre_encoder`` - written every ``save_interval`` epochs, newest one resumed: train.py:339-378, 627-637; like the
reference, discriminator weights are not checkpointed, App. B13).

The arithmetic of an iteration is ``mqgan_b200.training.TrainStep`` (CUDA only).  Launched under ``torchrun`` it
trains data-parallel: rank r takes batches r, r + N, ...; gradients are averaged inside the step.

Not carried over: wandb logging / image plots (losses are printed and appended to ``train_log.jsonl``), and dropout
(the step implements dropout 0; a config with generator.dropout > 0 is accepted and the value ignored with a
warning, because the reference's hard-wired 0.1 layers make "the configured dropout" ill-defined anyway).
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import random
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import spec as S


class RealMelSpectrogramDataset(torch.utils.data.Dataset):
    """Recursive ``*.npy`` mels of shape (T, n_mels) (train.py:84-134).  ``crop_len`` None = full length."""

    def __init__(self, real_dir: str, crop_len: Optional[int] = None):
        self.real_dir, self.crop_len = real_dir, crop_len
        self.filenames = sorted(os.path.join(root, fn) for root, _, files in os.walk(real_dir) for fn in files
                                if fn.endswith(".npy"))
        print(f"Found {len(self.filenames)} .npy files." if self.filenames
              else f"Warning: No .npy files found in {real_dir} (recursively).")

    def __len__(self) -> int:
        return len(self.filenames)

    def __getitem__(self, idx):
        path = self.filenames[idx]
        try:
            mel = np.load(path)
        except Exception as e:                                   # the reference skips unreadable files (train.py:110-113)
            print(f"[Load error] {path}: {e}")
            return None
        if mel.ndim != 2:
            print(f"[Shape error] {path}: shape={mel.shape}")
            return None
        full_len = mel.shape[0]
        target = self.crop_len
        if target is not None:
            if full_len > target:
                start = np.random.randint(0, full_len - target + 1)
                mel = mel[start:start + target]
            elif full_len < target:
                mel = np.concatenate([mel, np.zeros((target - full_len, mel.shape[1]), dtype=mel.dtype)], axis=0)
        mel_len = min(full_len, target) if target is not None else full_len
        return torch.as_tensor(mel.astype(np.float32)), int(mel_len), os.path.basename(path)


def pad_collate_fn(batch, crop_lens=None):
    """train.py:138-198: drop failed loads; ``crop_lens`` None = pad to the batch maximum, an int = crop / pad to it,
    a list = one length drawn per BATCH; longer items are randomly cropped, shorter ones zero-padded on the right and
    keep their true length."""
    batch = [item for item in batch if item is not None]
    if not batch:
        return None, None, None
    mels, lens, names = zip(*batch)
    tgt = None
    if crop_lens is not None:
        tgt = int(random.choice(crop_lens)) if isinstance(crop_lens, (list, tuple)) else int(crop_lens)
    if tgt is None:
        tgt = max(lens)
    out, new_lens = [], []
    for mel, full_len in zip(mels, lens):
        mel = torch.as_tensor(mel, dtype=torch.float32)
        if full_len > tgt:
            start = random.randint(0, full_len - tgt)
            mel = mel[start:start + tgt]
        elif mel.shape[0] < tgt:
            mel = torch.nn.functional.pad(mel, (0, 0, 0, tgt - mel.shape[0]))
        out.append(mel[:tgt])
        new_lens.append(min(full_len, tgt))
    return torch.stack(out), torch.tensor(new_lens, dtype=torch.int32), names


def split_dataset(n: int, validation_split: float, seed: int) -> Tuple[List[int], List[int]]:
    """train.py:249-257: ``random_split`` with a generator seeded by training.seed."""
    eval_size = int(validation_split * n)
    train_size = n - eval_size
    if train_size <= 0 or eval_size < 0:
        raise ValueError(f"Invalid train/eval split sizes. Train: {train_size}, Eval: {eval_size}.")
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(seed)).tolist()
    return perm[:train_size], perm[train_size:]


def epoch_schedule(train_idx: Sequence[int], batch_size: int, seed: int, epoch: int, rank: int, world: int, crop_len):
    """[(dataset indices, crop length)] of one replica for one epoch.  The shuffle and the crop-length draws come from ONE
    random stream seeded by (seed, epoch), identical on every replica: all replicas see the same crop length in the same
    iteration (the reference draws it per batch, train.py:153-158), so they capture / replay the same CUDA graphs and
    issue their gradient all-reduces in lock-step; rank r takes batches r, r + world, ...; a trailing group of fewer than
    ``world`` batches is dropped so every replica runs the same number of iterations."""
    order = list(train_idx)
    shared = random.Random(seed * 1000003 + epoch)
    shared.shuffle(order)
    chunks = [order[i:i + batch_size] for i in range(0, len(order), batch_size)]
    usable = len(chunks) - (len(chunks) % world if world > 1 else 0)
    out = []
    for i0 in range(0, usable, world):
        tgt = shared.choice(list(crop_len)) if isinstance(crop_len, (list, tuple)) else crop_len
        out.append((chunks[i0 + rank], tgt))
    return out


def latest_checkpoint(output_dir: str) -> Optional[str]:
    return max(glob.glob(os.path.join(output_dir, "checkpoint_epoch_*.pth")), key=os.path.getctime, default=None)


def build_arg_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Train an MQGAN model.")
    p.add_argument("--config", type=str, default="config.yaml", help="Path to the configuration file.")
    p.add_argument("--pretrained", type=str, default=None, help="Path to a pretrained checkpoint to load.")
    p.add_argument("--output_dir", type=str, default=None, help="Path to the output directory, overriding config.")
    # additive options (not in the reference)
    p.add_argument("--max_steps", type=int, default=None, help="stop after this many iterations (smoke runs)")
    p.add_argument("--no_graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    return p


class Trainer:
    """train.py:202-652 around ``TrainStep``."""

    def __init__(self, config: dict, max_steps: Optional[int] = None, use_graph: bool = True):
        import torch.distributed as dist
        from .preencoder import PreEncoder
        from .synth import synth_disc_state_dict
        from .training import TrainStep
        self.config, self.max_steps, self.use_graph = config, max_steps, use_graph
        if not torch.cuda.is_available() or config["training"].get("no_cuda", False):
            raise RuntimeError("mqgan_b200 training needs a CUDA device (there is no CPU path)")
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        self.device = torch.device("cuda", local)
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=self.device)
        seed = int(config["training"]["seed"])
        random.seed(seed + self.rank)
        np.random.seed(seed + self.rank)
        torch.manual_seed(seed)                                        # same initial weights on every replica
        self.start_epoch = 1
        data = config["data"]
        self.dataset = RealMelSpectrogramDataset(data["data_dir"], None)
        if len(self.dataset) == 0:
            raise ValueError("Dataset is empty after initialization.")
        self.train_idx, self.eval_idx = split_dataset(len(self.dataset), data["validation_split"], seed)
        mc = config["model"]
        gen = mc["generator"]
        if float(gen.get("dropout", 0.0)) != 0.0:
            print(f"Warning: generator.dropout = {gen['dropout']} ignored; this training step implements dropout 0 "
                  "(recorded as effective_dropout: 0.0 in every checkpoint's config).")
        # the regularisation actually applied, so a checkpoint says what produced it (the reference also hard-wires
        # p = 0.1 layers, preencoder.py:109,121,233, which this step does not apply either)
        config.setdefault("training", {})["effective_dropout"] = 0.0
        self.cfg = S.PreEncoderConfig.from_yaml_dict(config)
        self.pd_cfg = S.PatchDiscConfig.from_patch_yaml(self.cfg.mel_channels, mc["discriminator_patch"])
        self.mb_cfg = S.MultiBinConfig.from_yaml(self.cfg.mel_channels, mc["discriminator_multibin"])
        g_sd = PreEncoder(self.cfg.mel_channels, list(self.cfg.channels), list(self.cfg.kernel_sizes),
                          fsq_levels=list(self.cfg.fsq_levels), dropout=0.0,
                          refiner_base_channels=self.cfg.refiner_base_channels, refiner_depth=self.cfg.refiner_depth,
                          refiner_hidden_proj_divisor=self.cfg.refiner_hidden_proj_divisor).state_dict()
        os.makedirs(data["output_dir"], exist_ok=True)
        ckpt = latest_checkpoint(data["output_dir"])
        pretrained = config["training"].get("pretrained")
        resume = None
        if ckpt:
            print(f"=> Loading full checkpoint for resuming training from '{ckpt}'")
            resume = torch.load(ckpt, map_location="cpu", weights_only=False)
            g_sd = {k.replace("module.", ""): v for k, v in resume["model_state_dict"].items()}
            self.start_epoch = int(resume["epoch"]) + 1
        elif pretrained and os.path.isfile(pretrained):
            print(f"=> Loading pretrained generator from '{pretrained}'")
            sd = torch.load(pretrained, map_location="cpu", weights_only=False)
            sd = sd.get("model_state_dict", sd)
            g_sd.update({k.replace("module.", ""): v for k, v in sd.items() if k.replace("module.", "") in g_sd})
        else:
            print("No pretrained checkpoint specified or found. Training from scratch.")
        # discriminators: N(0, 0.02) conv weights as discriminators.py:193-198 (never checkpointed, App. B13)
        # with zero conv biases as discriminators.py:177-181 / :199-200 initialise them
        pd_sd = synth_disc_state_dict(S.patch_disc_param_spec(self.pd_cfg), seed=seed, zero_bias=True)
        mb_sd = synth_disc_state_dict(S.multibin_param_spec(self.mb_cfg), seed=seed + 1, zero_bias=True)
        tcfg = dict(S.TRAIN_DEFAULTS)
        tcfg.update(config["training"])
        self.step = TrainStep(self.cfg, self.pd_cfg, self.mb_cfg, g_sd, pd_sd, mb_sd, tcfg, self.device, d_autocast_bf16=True)
        if resume is not None:
            for name, opt in (("optimizer_g_state_dict", self.step.opt_g), ("optimizer_d_state_dict", self.step.opt_d)):
                try:
                    opt.load_state_dict(resume[name])
                except Exception as e:                                 # a reference-written optimiser state (different layout)
                    print(f"Warning: could not restore {name}: {e}")
            # load_state_dict replaces param_groups[*]['lr'] with a deep-copied CPU tensor: re-bind the live device
            # tensors the step writes the warm-up schedule into, or the LR would stay frozen at the checkpoint's value
            self.step.rebind_lr()
            self.step.g_steps = int(resume.get("g_steps", 0))
        self.iterations = 0

    def batches(self, epoch: int):
        """Shuffled training batches of this epoch; rank r takes batches r, r + world, ..."""
        for idx, tgt in epoch_schedule(self.train_idx, int(self.config["data"]["batch_size"]), int(self.config["training"]["seed"]),
                                       epoch, self.rank, self.world, self.config["data"].get("crop_len")):
            yield pad_collate_fn([self.dataset[j] for j in idx], crop_lens=tgt)

    def train_epoch(self, epoch: int) -> Optional[dict]:
        self.step.start_epoch()                                        # train.py:504-506
        gan = epoch >= int(self.config["training"]["discriminator_train_start_epoch"])
        last = None
        for real, lens, _ in self.batches(epoch):
            have = real is not None and real.size(0) > 0
            if self.world > 1:
                import torch.distributed as dist
                # the step issues collectives (LeCam means, gradient buckets): skipping must be unanimous or the other
                # ranks block in their all-reduces
                flag = torch.tensor([1 if have else 0], device=self.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                have = bool(int(flag.item()))
            if not have:
                continue
            key = (tuple(real.shape), bool(gan), None)
            # graphs are captured per batch shape once the discriminators have left their first (training-mode) iteration
            if self.use_graph and not self.step.d_training and real.size(0) == int(self.config["data"]["batch_size"]):
                if key not in self.step._graphs:                       # capture() trains on this batch once (its warm-up)
                    self.step.capture(real, lens.long(), gan=gan, warmup=1)
                    out = self.step.last_losses
                else:
                    out = self.step.step_graphed(real, lens.long(), gan=gan)
            else:
                out = self.step.step(real, lens.long(), gan=gan)
            self.iterations += 1
            last = {k: float(v) for k, v in out.items()}
            last.update(epoch=epoch, iteration=self.iterations, learning_rate=self.step.current_lr_g())
            if self.rank == 0:
                with open(os.path.join(self.config["data"]["output_dir"], "train_log.jsonl"), "a") as f:
                    f.write(json.dumps(last) + "\n")
            if self.max_steps is not None and self.iterations >= self.max_steps:
                break
        if last is not None and self.rank == 0:
            print(f"Epoch [{epoch}/{self.config['training']['num_epochs']}] D_loss={last['loss_d']:.4f} "
                  f"G_loss={last['loss_g_total']:.4f} Recon_Post={last['loss_recon_post']:.4f}")
        return last

    def save_checkpoint(self, epoch: int) -> str:
        """train.py:627-637 (GradScaler states are empty dicts: bf16 needs no loss scaling)."""
        path = os.path.join(self.config["data"]["output_dir"], f"checkpoint_epoch_{epoch:03d}.pth")
        torch.save({"epoch": epoch,
                    "model_state_dict": {k: v.detach().cpu() for k, v in self.step.generator_state_dict().items()},
                    "optimizer_g_state_dict": self.step.opt_g.state_dict(),
                    "optimizer_d_state_dict": self.step.opt_d.state_dict(),
                    "scaler_g_state_dict": {}, "scaler_d_state_dict": {}, "g_steps": self.step.g_steps,
                    "config": self.config}, path)
        print(f"Checkpoint saved to {path}")
        return path

    def train(self):
        log = self.config.get("logging", {})
        for epoch in range(self.start_epoch, int(self.config["training"]["num_epochs"]) + 1):
            self.train_epoch(epoch)
            done = self.max_steps is not None and self.iterations >= self.max_steps
            if self.rank == 0 and (epoch % int(log.get("save_interval", 1)) == 0 or done):
                self.save_checkpoint(epoch)
            if done:
                break
        print("Training finished.")


def main(argv: Optional[Sequence[str]] = None) -> int:
    import yaml
    args = build_arg_parser().parse_args(argv)
    with open(args.config, "r") as f:
        config = yaml.safe_load(f)
    if args.pretrained:
        config["training"]["pretrained"] = args.pretrained
    if args.output_dir:
        config["data"]["output_dir"] = args.output_dir
    Trainer(config, max_steps=args.max_steps, use_graph=not args.no_graph).train()
    return 0
