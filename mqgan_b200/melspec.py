"""Log-mel front-end on the GPU: the step before the re-encode path.

Reference: convert_spectrograms.py:14-35 - torchaudio ``MelSpectrogram(power=1)`` with its default
hann window / center=True / reflect padding / htk mel scale / norm=None, then
``log(clamp(min=1e-5))``, returned as (frames, n_mels).  ``LogMelExtractor`` prepares the constant
tables once (window, FFT twiddles, sparse mel filterbank) and runs ``mq_log_mel`` (csrc/melspec.cu):
one fused kernel, waveform in -> log-mel out.  There is no CPU or torchaudio fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import MelspecParams
from .ops import on_device

REQUIRED_KEYS = ("sampling_rate", "filter_length", "hop_length", "win_length", "n_mel_channels", "mel_fmin", "mel_fmax")


def _hann_padded(win_length: int, n_fft: int) -> np.ndarray:
    n = np.arange(win_length, dtype=np.float64)
    w = 0.5 - 0.5 * np.cos(2.0 * math.pi * n / win_length)          # torch.hann_window(periodic=True)
    out = np.zeros(n_fft, dtype=np.float64)
    left = (n_fft - win_length) // 2                                  # torch.stft centres a short window
    out[left:left + win_length] = w
    return out


def _mel_fbanks(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> np.ndarray:
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale="htk") -> (n_freqs, n_mels), float64."""
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    f_pts = 700.0 * (10.0 ** (np.linspace(m_min, m_max, n_mels + 2) / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    return np.maximum(0.0, np.minimum(-slopes[:, :-2] / f_diff[:-1], slopes[:, 2:] / f_diff[1:]))


class LogMelExtractor:
    """spec: the ``spectrogram`` section of the reference's spec_config_*.yaml."""

    def __init__(self, spec: Dict, device="cuda", clip_val: float = 1e-5):
        for k in REQUIRED_KEYS:
            if k not in spec:
                raise ValueError(f"Missing required key in config['spectrogram']: '{k}'")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("LogMelExtractor needs a CUDA device (no CPU fallback)")
        self.n_fft, self.hop, self.win = int(spec["filter_length"]), int(spec["hop_length"]), int(spec["win_length"])
        self.n_mels, self.sr = int(spec["n_mel_channels"]), int(spec["sampling_rate"])
        if self.n_fft & (self.n_fft - 1) or not 64 <= self.n_fft <= 4096:
            raise ValueError("filter_length must be a power of two in [64, 4096]")
        if self.win > self.n_fft:
            raise ValueError("win_length must not exceed filter_length")
        self.n_freqs = self.n_fft // 2 + 1
        self.clip = float(clip_val)
        dev = self.device
        self.window = torch.from_numpy(_hann_padded(self.win, self.n_fft).astype(np.float32)).to(dev)
        t = np.arange(self.n_fft, dtype=np.float64) * (2.0 * math.pi / self.n_fft)
        self.twiddle = torch.from_numpy(np.stack([np.cos(t), -np.sin(t)], axis=1).astype(np.float32)).contiguous().to(dev)
        fb = _mel_fbanks(self.n_freqs, float(spec["mel_fmin"]), float(spec["mel_fmax"]), self.n_mels, self.sr).astype(np.float32)
        start, count, off, w = [], [], [], []
        for m in range(self.n_mels):
            nz = np.nonzero(fb[:, m])[0]
            s, c = (int(nz[0]), int(nz[-1] - nz[0] + 1)) if len(nz) else (0, 0)
            start.append(s); count.append(c); off.append(len(w)); w.extend(fb[s:s + c, m].tolist())
        self.fb_start = torch.tensor(start, dtype=torch.int32, device=dev)
        self.fb_count = torch.tensor(count, dtype=torch.int32, device=dev)
        self.fb_off = torch.tensor(off, dtype=torch.int32, device=dev)
        self.fb_w = torch.tensor(w if w else [0.0], dtype=torch.float32, device=dev)

    def num_frames(self, n_samples: int) -> int:
        return 1 + n_samples // self.hop if n_samples > self.n_fft // 2 else 0

    @on_device
    def __call__(self, wav: torch.Tensor, lengths: Optional[Sequence[int]] = None) -> Tuple[torch.Tensor, List[int]]:
        """wav (B, Tmax) fp32 (zero-padded past each length) -> (log-mel (B, Fmax, n_mels) on the device,
        frames per utterance).  Rows past an utterance's frame count are zero."""
        if wav.dim() != 2:
            raise ValueError(f"wav must be (B, T), got {tuple(wav.shape)}")
        wav = wav.to(self.device, torch.float32).contiguous()
        B, Tmax = wav.shape
        lens = [Tmax] * B if lengths is None else [int(x) for x in lengths]
        if len(lens) != B or any(l < 0 or l > Tmax for l in lens):
            raise ValueError("lengths must hold one value in [0, T] per row")
        frames = [self.num_frames(l) for l in lens]
        fmax = max(max(frames), 1)
        out = torch.empty(B, fmax, self.n_mels, dtype=torch.float32, device=self.device)
        ld = torch.tensor(lens, dtype=torch.int64, device=self.device)
        p = MelspecParams()
        p.wav, p.wav_ld, p.lengths, p.B = wav.data_ptr(), Tmax, ld.data_ptr(), B
        p.n_fft, p.hop, p.n_mels, p.n_freqs = self.n_fft, self.hop, self.n_mels, self.n_freqs
        p.window, p.twiddle = self.window.data_ptr(), self.twiddle.data_ptr()
        p.fb_start, p.fb_count, p.fb_off, p.fb_w = (self.fb_start.data_ptr(), self.fb_count.data_ptr(),
                                                    self.fb_off.data_ptr(), self.fb_w.data_ptr())
        p.clip, p.out, p.out_frames = self.clip, out.data_ptr(), fmax
        _lib.call("mq_log_mel", C.byref(p), torch.cuda.current_stream().cuda_stream)
        return out, frames
