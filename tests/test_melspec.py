"""Log-mel front-end: oracle vs the reference's own outputs (CPU), CUDA kernel vs both (GPU)."""
import os

import numpy as np
import pytest
import torch
import yaml

from oracle import mel_oracle as M

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["mel_hifispeech", "mel_hifispeech_short", "mel_hifimusic"]


def _load(name):
    fx = np.load(os.path.join(GOLDEN, name + ".npz"))
    spec = yaml.safe_load(str(fx["spec_yaml"]))
    wav = M.synth_wave(int(fx["seed"]), int(fx["n_samples"]), int(spec["sampling_rate"]))
    return spec, wav, fx["mel"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_extractor_outputs(name):
    """The restatement reproduces convert_spectrograms.TorchMelSpectrogramExtractor (torchaudio 2.11.0) to the
    reference's own fp32 noise: log-mel within 1e-4 everywhere."""
    spec, wav, ref = _load(name)
    out = M.log_mel(wav, spec, np.float64)
    assert out.shape == ref.shape == (1 + len(wav) // spec["hop_length"], spec["n_mel_channels"])
    assert np.abs(out - ref).max() < 1e-4


def test_filterbank_and_window_definitions():
    fb = M.mel_filterbank(1025, 0.0, 22050.0, 128, 44100)
    assert fb.shape == (1025, 128) and fb.min() >= 0 and fb.max() <= 1.0
    assert (np.count_nonzero(fb, axis=0) > 0).all()           # every mel bin has support at n_fft = 2048
    assert np.count_nonzero(fb) < 3 * 1025                     # triangular: a frequency feeds at most two bins (+ edges)
    w = M.hann_window_padded(2048, 2048)
    assert w[0] == 0.0 and abs(w[1024] - 1.0) < 1e-12          # periodic hann
    w2 = M.hann_window_padded(1024, 2048)
    assert w2[:512].max() == 0.0 and w2[512 + 512] == pytest.approx(1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_log_mel_vs_reference_and_oracle(name):
    from mqgan_b200.convert_spectrograms import TorchMelSpectrogramExtractor
    spec, wav, ref = _load(name)
    ext = TorchMelSpectrogramExtractor(spec, device="cuda")
    out = ext.get_mel_from_wav(torch.from_numpy(wav)[None, :])
    assert not out.is_cuda and out.dtype == torch.float32 and tuple(out.shape) == ref.shape
    o64 = M.log_mel(wav, spec, np.float64)
    err_ref = float(np.abs(out.numpy() - ref).max())
    err_64 = float(np.abs(out.numpy() - o64).max())
    ref_64 = float(np.abs(ref - o64).max())
    print(name, "cuda vs reference", err_ref, "cuda vs float64", err_64, "reference vs float64", ref_64)
    # fp32 FFT + fp32 mel sums: the same noise level as the reference's own fp32 path (2-3e-5 here)
    assert err_64 < 1e-4 and err_ref < 1.5e-4


@pytest.mark.gpu
def test_cuda_log_mel_batched_ragged_and_edges():
    from mqgan_b200.melspec import LogMelExtractor
    spec = {"sampling_rate": 16000, "filter_length": 512, "hop_length": 160, "win_length": 400,
            "n_mel_channels": 40, "mel_fmin": 50.0, "mel_fmax": 7600.0}
    ext = LogMelExtractor(spec, device="cuda")
    lens = [4000, 3999, 257, 256, 1601]                      # 256 = n_fft/2: no frames (torch.stft rejects it)
    wavs = [M.synth_wave(20 + i, max(l, 1), 16000)[:l] for i, l in enumerate(lens)]
    batch = torch.zeros(len(lens), max(lens))
    for i, w in enumerate(wavs):
        batch[i, : len(w)] = torch.from_numpy(w)
    out, frames = ext(batch.cuda(), lens)
    assert frames == [26, 25, 2, 0, 11]
    out = out.cpu().numpy()
    for i, (w, f) in enumerate(zip(wavs, frames)):
        if f:
            ref = M.log_mel(w, spec, np.float64)
            assert ref.shape[0] == f
            assert np.abs(out[i, :f] - ref).max() < 1e-4, i
        assert np.all(out[i, f:] == 0.0)                     # rows past the utterance's frames are zero
    # an odd number of frames exercises the lone real-only FFT of the last CTA; a short window is centred
    single, fr = ext(torch.from_numpy(wavs[0])[None, :].cuda())
    assert fr == [26] and np.abs(single[0].cpu().numpy() - out[0, :26]).max() == 0.0
    with pytest.raises(ValueError):
        ext(torch.zeros(4000).cuda())
