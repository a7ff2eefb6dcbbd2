"""Generate tests/golden/mel_*.npz by running the reference's own mel extractor
(convert_spectrograms.py:14-35, torchaudio) in the build container:

    python oracle/make_golden_mel.py

Inputs are regenerated anywhere from (seed, n_samples) by oracle.mel_oracle.synth_wave; only the
reference's OUTPUTS are stored."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import yaml

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
REF = os.environ.get("MQGAN_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from oracle.mel_oracle import synth_wave  # noqa: E402

CASES = [  # name, spec config, seed, samples
    ("mel_hifispeech", "spec_config_hifispeech.yaml", 1, 50000),
    ("mel_hifispeech_short", "spec_config_hifispeech.yaml", 2, 44100 + 17),
    ("mel_hifimusic", "spec_config_hifimusic.yaml", 3, 70001),
]


def main():
    import convert_spectrograms as ref_cs                     # the unmodified reference module
    out_dir = os.path.join(REPO, "tests", "golden")
    for name, cfg_file, seed, n in CASES:
        with open(os.path.join(REF, "configs", cfg_file)) as f:
            spec = yaml.safe_load(f)["spectrogram"]
        wav = synth_wave(seed, n, int(spec["sampling_rate"]))
        ext = ref_cs.TorchMelSpectrogramExtractor(spec)
        with torch.no_grad():
            mel = ext.get_mel_from_wav(torch.from_numpy(wav)[None, :]).contiguous().numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), mel=mel.astype(np.float32), seed=seed, n_samples=n,
                            spec_yaml=yaml.safe_dump(spec), torchaudio=str(__import__("torchaudio").__version__))
        print(name, mel.shape, float(mel.min()), float(mel.max()))


if __name__ == "__main__":
    main()
