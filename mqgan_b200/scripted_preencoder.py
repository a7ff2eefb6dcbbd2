"""Drop-in ``ScriptedPreEncoder`` (reference: scripted_preencoder.py:24-175).

Loads the directory the reference's convert_to_torchscript.py writes
(``model_cuda.pt`` / ``model_cpu.pt`` + ``model_config.yaml``), but only harvests the
TorchScript module's ``state_dict()`` (142 tensors) and the architecture
from the YAML; ``encode`` / ``decode`` then run on the sm_100a kernels.  Because the
weights are device-agnostic, a directory holding only ``model_cpu.pt`` still runs on
the GPU (the reference would silently fall back to CPU, :84-87).  A CPU device is
refused: this build has no CPU path.
"""
from __future__ import annotations

import os
from typing import List, Optional, Union

import torch
import yaml

from .preencoder import PreEncoder, _accept_stripped_weight_norm, sequence_mask  # noqa: F401
from .spec import PreEncoderConfig


class ScriptedPreEncoder:
    def __init__(self, model_dir: str, device: Optional[str] = "cpu"):
        if not os.path.isdir(model_dir):
            raise FileNotFoundError(f"Model directory not found: {model_dir}")
        self.device = torch.device(device)
        config_path = os.path.join(model_dir, "model_config.yaml")
        if not os.path.exists(config_path):
            raise FileNotFoundError(f"model_config.yaml not found in: {model_dir}")
        try:
            with open(config_path, "r") as f:
                self.config = yaml.safe_load(f)
        except Exception as e:
            raise RuntimeError(f"Failed to load or parse config file: {e}")
        if self.device.type != "cuda":
            raise RuntimeError("mqgan_b200.ScriptedPreEncoder runs on CUDA (B200) only; pass device='cuda' "
                               "(there is no CPU fallback in this build)")
        model_path = self._get_model_path(model_dir)
        try:
            scripted = torch.jit.load(model_path, map_location="cpu")
            weights = {k: v.detach().clone() for k, v in scripted.state_dict().items()}
            cfg = PreEncoderConfig.from_yaml_dict(self.config)
            model = PreEncoder(cfg.mel_channels, list(cfg.channels), list(cfg.kernel_sizes),
                               fsq_levels=list(cfg.fsq_levels), dropout=0.0,
                               refiner_base_channels=cfg.refiner_base_channels, refiner_depth=cfg.refiner_depth,
                               refiner_hidden_proj_divisor=cfg.refiner_hidden_proj_divisor)
            model.load_state_dict(_accept_stripped_weight_norm(model, weights), strict=True)
            self.model = model.to(self.device).eval()
            print(f"Successfully loaded model from {os.path.basename(model_path)} onto {self.device}.")
        except Exception as e:
            raise RuntimeError(f"Failed to load TorchScript model from {model_path}: {e}")

    def _get_model_path(self, model_dir: str) -> str:
        cuda_path = os.path.join(model_dir, "model_cuda.pt")
        cpu_path = os.path.join(model_dir, "model_cpu.pt")
        if os.path.exists(cuda_path):
            return cuda_path
        if os.path.exists(cpu_path):
            return cpu_path          # weights only: still executed on the GPU
        raise FileNotFoundError("No CUDA or CPU model found in the specified directory.")

    @property
    def mel_channels(self) -> int:
        return self.config.get("model", {}).get("mel_channels", 0)

    @property
    def fsq_levels(self) -> List[int]:
        return self.config.get("model", {}).get("generator", {}).get("fsq_levels", [])

    def _prepare_mask(self, max_len: int, lengths: torch.Tensor) -> torch.Tensor:
        return sequence_mask(max_len, lengths.to(self.device)).unsqueeze(1)

    def _mask(self, T, lengths):
        if lengths is None:
            return None
        if not isinstance(lengths, torch.Tensor):
            lengths = torch.tensor(lengths, dtype=torch.long)
        return self._prepare_mask(T, lengths)

    def encode(self, spectrogram: torch.Tensor, lengths: Optional[Union[List[int], torch.Tensor]] = None) -> torch.Tensor:
        if spectrogram.ndim != 3:
            raise ValueError(f"Input spectrogram must be a 3D tensor (B, T, C), but got shape {spectrogram.shape}")
        spectrogram = spectrogram.to(self.device)
        mask = self._mask(spectrogram.shape[1], lengths)
        with torch.no_grad():
            try:
                return self.model.encode(spectrogram, mask)
            except Exception as e:
                raise RuntimeError(f"An error occurred during the encode operation: {e}")

    def decode(self, indices: torch.Tensor, lengths: Optional[Union[List[int], torch.Tensor]] = None) -> torch.Tensor:
        indices = indices.to(self.device)
        mask = self._mask(indices.shape[1], lengths)
        with torch.no_grad():
            try:
                # host-side lengths let the engine skip most of a ragged batch's padding (identical output)
                host_lens = None if (lengths is None or (isinstance(lengths, torch.Tensor) and lengths.is_cuda)) else lengths
                return self.model.decode(indices, mask, lengths=host_lens)
            except Exception as e:
                raise RuntimeError(f"An error occurred during the decode operation: {e}")
