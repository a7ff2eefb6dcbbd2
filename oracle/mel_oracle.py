"""TEST INFRASTRUCTURE - CPU restatement of the reference's mel front-end.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; nothing
under mqgan_b200/ does.

Reference: convert_spectrograms.py:14-35 (`TorchMelSpectrogramExtractor`): torchaudio
`MelSpectrogram(sample_rate, n_fft, win_length, hop_length, n_mels, f_min, f_max, power=1.0)` followed by
`log(clamp(mel, min=1e-5))`, returned as (frames, n_mels).  The arithmetic lives in the third-party
dependency torchaudio (requirements: unpinned; this container has 2.11.0), whose published algorithm
with the defaults the reference leaves untouched is restated here:

  Spectrogram (torchaudio/transforms/_transforms.py `Spectrogram`, functional.spectrogram):
      window = hann_window(win_length, periodic=True), zero-padded on both sides to n_fft;
      center=True, pad_mode="reflect": the signal is reflect-padded by n_fft//2 on each side;
      frames = 1 + T // hop; S[f, k] = | sum_n x_pad[f*hop + n] * window[n] * exp(-2 pi i k n / n_fft) |,
      k = 0 .. n_fft/2 (onesided), no normalisation, power = 1 (magnitude).
  MelScale (functional.melscale_fbanks, mel_scale="htk", norm=None):
      m(f) = 2595 log10(1 + f/700); n_mels + 2 points equally spaced in mel between f_min and f_max;
      triangular filters over all_freqs = linspace(0, sample_rate // 2, n_freqs); fb (n_freqs, n_mels);
      mel = fb^T S.

Pinned by tests/golden/mel_*.npz, outputs of the reference class itself run in the build container
(oracle/make_golden_mel.py).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np


def hann_window_padded(win_length: int, n_fft: int, dtype=np.float64) -> np.ndarray:
    """torch.hann_window(win_length, periodic=True), centred in n_fft (torch.stft pads the window)."""
    n = np.arange(win_length, dtype=np.float64)
    w = 0.5 - 0.5 * np.cos(2.0 * math.pi * n / win_length)
    left = (n_fft - win_length) // 2
    out = np.zeros(n_fft, dtype=np.float64)
    out[left:left + win_length] = w
    return out.astype(dtype)


def mel_filterbank(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int,
                   dtype=np.float64) -> np.ndarray:
    """torchaudio.functional.melscale_fbanks(..., norm=None, mel_scale="htk") -> (n_freqs, n_mels)."""
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = np.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]                       # (n_mels + 1)
    slopes = f_pts[None, :] - all_freqs[:, None]          # (n_freqs, n_mels + 2)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(0.0, np.minimum(down, up)).astype(dtype)


def log_mel(wav: np.ndarray, spec: Dict, dtype=np.float64, clip: float = 1e-5) -> np.ndarray:
    """wav (T,) -> (frames, n_mels) log-mel, convert_spectrograms.py:31-35."""
    n_fft, hop, win = int(spec["filter_length"]), int(spec["hop_length"]), int(spec["win_length"])
    x = np.asarray(wav, dtype=dtype).reshape(-1)
    T = x.shape[0]
    if T <= n_fft // 2:
        raise ValueError("reflect padding needs more than n_fft/2 samples")
    xp = np.pad(x, (n_fft // 2, n_fft // 2), mode="reflect")
    frames = 1 + T // hop
    idx = np.arange(frames)[:, None] * hop + np.arange(n_fft)[None, :]
    window = hann_window_padded(win, n_fft, dtype)
    S = np.abs(np.fft.rfft((xp[idx] * window[None, :]).astype(dtype), axis=1))            # (frames, n_freqs)
    fb = mel_filterbank(n_fft // 2 + 1, float(spec["mel_fmin"]), float(spec["mel_fmax"]), int(spec["n_mel_channels"]),
                        int(spec["sampling_rate"]), dtype)
    mel = S.astype(dtype) @ fb
    return np.log(np.maximum(mel, dtype(clip) if dtype is not np.float64 else clip)).astype(dtype)


def synth_wave(seed: int, n_samples: int, sample_rate: int = 44100) -> np.ndarray:
    """Deterministic test signal: a few sines with slow amplitude modulation plus noise, |x| < 1."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples) / sample_rate
    x = 0.05 * rng.standard_normal(n_samples)
    for _ in range(6):
        f = float(rng.uniform(80.0, 9000.0))
        a = float(rng.uniform(0.02, 0.15))
        x += a * np.sin(2 * math.pi * f * t + rng.uniform(0, 6.28)) * (0.6 + 0.4 * np.sin(2 * math.pi * rng.uniform(0.5, 4.0) * t))
    return np.clip(x, -0.99, 0.99).astype(np.float32)
