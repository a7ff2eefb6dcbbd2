"""Drop-in for the reference's convert_spectrograms.py (audio -> log-mel .npy), same class and
function names, same config keys and CLI flags; the mel extraction runs in ``mq_log_mel`` on the GPU.

Reference: convert_spectrograms.py:14-35 (TorchMelSpectrogramExtractor), :37-66
(MelSpectrogramConverter), :67-90 (worker / chunkify / validate_config), :91-133 (main).
Audio decoding / resampling stay on the host with torchaudio, as in the reference (that is file I/O,
not the path).  Differences, all additive: files are processed by one worker per visible GPU instead
of one per CPU core, and ``TorchMelSpectrogramExtractor.get_mel_from_wav`` accepts a CUDA tensor.
"""
from __future__ import annotations

import argparse
import os
from typing import Dict

import numpy as np
import torch

from .melspec import LogMelExtractor


class TorchMelSpectrogramExtractor:
    def __init__(self, spec_config: Dict, device="cuda"):
        self.config = spec_config
        self.clip_val = 1e-5
        self.transf = LogMelExtractor(spec_config, device=device, clip_val=self.clip_val)

    def get_mel_from_wav(self, wav: torch.Tensor) -> torch.Tensor:
        """wav: (1, T) -> mel_norm: (T_frames, n_mels), on the input's device (convert_spectrograms.py:31-35)."""
        if wav.dim() != 2 or wav.shape[0] != 1:
            raise ValueError(f"expected a (1, T) waveform, got {tuple(wav.shape)}")
        if wav.shape[1] <= self.transf.n_fft // 2:
            raise RuntimeError("reflect padding needs more than filter_length/2 samples")      # torch.stft raises here too
        out, frames = self.transf(wav)
        mel = out[0, : frames[0]]
        return mel if wav.is_cuda else mel.cpu()


class MelSpectrogramConverter:
    def __init__(self, config, device="cuda"):
        self.config = config
        self.extractor = TorchMelSpectrogramExtractor(config["spectrogram"], device=device)
        os.makedirs(self.config["io"]["output_folder"], exist_ok=True)

    def process_file(self, file_path, output_dir):
        base_name = os.path.splitext(os.path.basename(file_path))[0]
        output_file_path = os.path.join(output_dir, f"{base_name}_mel.npy")
        if os.path.isfile(output_file_path):
            return True
        try:
            import torchaudio
            wav_tensor, sr = torchaudio.load(file_path)
            target = self.config["spectrogram"]["sampling_rate"]
            if sr != target:
                wav_tensor = torchaudio.transforms.Resample(orig_freq=sr, new_freq=target)(wav_tensor)
            duration = wav_tensor.shape[1] / target
            if duration < 1.0 or duration > 15.0:
                return False
            # the reference feeds every channel through the transform and squeezes; like it, only mono survives
            mel_spectrogram = self.extractor.get_mel_from_wav(wav_tensor)
            np.save(output_file_path, mel_spectrogram.cpu().numpy())
            return True
        except Exception as e:  # noqa: BLE001 - the reference's catch-all (:63-65)
            print(f"Error processing {file_path}: {e}")
            return False


def worker(worker_id, tasks, config):
    if torch.cuda.is_available():
        torch.cuda.set_device(worker_id % torch.cuda.device_count())
    converter = MelSpectrogramConverter(config, device=f"cuda:{worker_id % max(torch.cuda.device_count(), 1)}")
    try:
        from tqdm import tqdm
        it = tqdm(tasks, desc=f"Worker {worker_id}", position=worker_id)
    except Exception:  # pragma: no cover
        it = tasks
    for file_path, output_dir in it:
        os.makedirs(output_dir, exist_ok=True)
        converter.process_file(file_path, output_dir)


def chunkify(lst, n):
    k, m = divmod(len(lst), n)
    return [lst[i * k + min(i, m):(i + 1) * k + min(i + 1, m)] for i in range(n)]


def validate_config(config):
    required_keys = {
        "io": ["input_folder", "output_folder", "audio_extensions"],
        "spectrogram": ["sampling_rate", "filter_length", "hop_length", "win_length", "n_mel_channels", "mel_fmin", "mel_fmax"],
    }
    for main_key, sub_keys in required_keys.items():
        if main_key not in config:
            raise ValueError(f"Missing required key in config: '{main_key}'")
        for sub_key in sub_keys:
            if sub_key not in config[main_key]:
                raise ValueError(f"Missing required key in config['{main_key}']: '{sub_key}'")


def collect_tasks(config):
    tasks = []
    audio_exts = tuple(config["io"]["audio_extensions"])
    for root, _, files in os.walk(config["io"]["input_folder"]):
        rel_path = os.path.relpath(root, config["io"]["input_folder"])
        output_subfolder = os.path.join(config["io"]["output_folder"], rel_path)
        for wav_file in files:
            if wav_file.lower().endswith(audio_exts):
                tasks.append((os.path.join(root, wav_file), output_subfolder))
    return tasks


def main(argv=None):
    import yaml
    parser = argparse.ArgumentParser(description="Convert audio files to mel spectrograms.")
    parser.add_argument("--config", type=str, default="spec_config.yaml", help="Path to the configuration file.")
    parser.add_argument("--input_folder", type=str, default=None, help="Override the input folder specified in the config file.")
    parser.add_argument("--output_folder", type=str, default=None, help="Override the output folder specified in the config file.")
    args = parser.parse_args(argv)
    with open(args.config, "r") as f:
        config = yaml.safe_load(f)
    if args.input_folder:
        config["io"]["input_folder"] = args.input_folder
    if args.output_folder:
        config["io"]["output_folder"] = args.output_folder
    try:
        validate_config(config)
    except ValueError as e:
        print(f"Configuration Error: {e}")
        raise SystemExit(1)
    os.makedirs(config["io"]["output_folder"], exist_ok=True)
    tasks = collect_tasks(config)
    num_workers = max(torch.cuda.device_count(), 1)
    if num_workers == 1:
        worker(0, tasks, config)
        return
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=worker, args=(i, chunk, config)) for i, chunk in enumerate(chunkify(tasks, num_workers))]
    for p in procs:
        p.start()
    for p in procs:
        p.join()


if __name__ == "__main__":
    main()
