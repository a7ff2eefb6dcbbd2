#!/usr/bin/env python
"""Drop-in for the reference's reencode_spectrograms_from_checkpoint.py (same flags;
reference: reencode_spectrograms_from_checkpoint.py:110-142), on the sm_100a kernels."""
import argparse
import functools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch  # noqa: E402
import yaml  # noqa: E402

from mqgan_b200 import reencode as R  # noqa: E402
from mqgan_b200.preencoder import get_pre_encoder, sequence_mask  # noqa: E402


class _CheckpointModel:
    """encode/decode(lengths=) adaptor over the raw PreEncoder (mask built as :77-81)."""

    def __init__(self, checkpoint_path, config, device):
        mp, gp = config['model'], config['model']['generator']
        self.device = torch.device(device)
        self.model = get_pre_encoder(
            model_path=checkpoint_path, device=device, mel_channels=mp['mel_channels'], channels=gp['channels'],
            kernel_sizes=gp['kernel_sizes'], fsq_levels=gp['fsq_levels'],
            refiner_base_channels=gp.get('refiner_base_channels', 128), refiner_depth=gp.get('refiner_depth', 3),
            refiner_hidden_proj_divisor=gp.get('refiner_hidden_proj_divisor', 8), inference=True)

    def _mask(self, T, lengths):
        lt = torch.tensor(lengths, dtype=torch.long, device=self.device)
        return sequence_mask(T, lt).unsqueeze(1)

    def encode(self, batch, lengths):
        batch = batch.float().to(self.device)
        with torch.no_grad():
            return self.model.encode(batch, x_mask=self._mask(batch.shape[1], lengths))

    def decode(self, idx, lengths):
        with torch.no_grad():
            return self.model.decode(idx, x_mask=self._mask(idx.shape[1], lengths), lengths=list(lengths))


def _make_model(checkpoint_path, config, device):
    return _CheckpointModel(checkpoint_path, config, device)


def reencode_spectrograms(checkpoint_path, config, input_dir, output_dir, device, batch_size, gpus=1,
                          sort_by_length=False):
    print(f"Loading model from checkpoint: {checkpoint_path}")
    rank, world, local = R.dist_env()
    if gpus > 1 and world == 1:
        done, failed = R.run_multi_gpu(functools.partial(_make_model, checkpoint_path, config), input_dir,
                                       output_dir, batch_size, gpus, sort_by_length)
    else:
        if world > 1 and torch.device(device).type == "cuda":
            device = f"cuda:{local}"
            torch.cuda.set_device(local)
        try:
            model = _CheckpointModel(checkpoint_path, config, device)
        except (FileNotFoundError, RuntimeError, KeyError) as e:
            print(f"Error: Could not load the model. {e}")
            return
        print(f"Searching for .npy files in: {input_dir}")

        def run(batch, lengths):
            return model.decode(model.encode(batch, lengths), lengths)

        import time
        t0 = time.time()
        done, failed = R.reencode_tree(run, input_dir, output_dir, batch_size, rank, world,
                                        sort_by_length=sort_by_length)
        if torch.device(device).type == "cuda":
            torch.cuda.synchronize()
        if os.environ.get("MQ_CLI_TIMING") == "1":
            import json
            print("MQ_CLI_TIMING " + json.dumps({"workers": world, "process_span_s": time.time() - t0,
                                                 "process_s_per_worker": [time.time() - t0]}), flush=True)
        done, failed = R.finish_distributed(done, failed)
    if rank == 0:
        print("\nProcessing complete.")
        print(f"Re-encoded {done} spectrograms ({failed} failed batches); saved to: {output_dir}")


def main():
    parser = argparse.ArgumentParser(
        description="Re-encode spectrograms using a raw PreEncoder model checkpoint (B200-native kernels).",
        formatter_class=argparse.RawTextHelpFormatter)
    parser.add_argument('--checkpoint', type=str, required=True,
                        help='Path to the raw PyTorch PreEncoder model checkpoint (.pth).')
    parser.add_argument('--config', type=str, required=True, help='Path to the model configuration YAML file.')
    parser.add_argument('--input_dir', type=str, required=True,
                        help='Path to the input folder containing .npy spectrograms.')
    parser.add_argument('--output_dir', type=str, required=True,
                        help='Path to the output folder where re-encoded spectrograms will be saved.')
    parser.add_argument('--device', type=str, default='cpu',
                        help='Device to use for inference (e.g., "cpu", "cuda"). This build only runs on "cuda".')
    parser.add_argument('--batch_size', type=int, default=32,
                        help='Number of spectrograms to process in a single batch. Defaults to 32.')
    parser.add_argument('--sort_by_length', action='store_true',
                        help='(added) batch files of similar length together (less padding, faster). Changes the '
                             'batch composition, which the encoder is sensitive to through padding: off by default.')
    parser.add_argument('--gpus', type=int, default=1, help='(added) shard batches over this many GPUs.')
    args = parser.parse_args()
    try:
        with open(args.config, 'r') as f:
            config = yaml.safe_load(f)
    except (FileNotFoundError, yaml.YAMLError) as e:
        print(f"Error loading config file: {e}")
        return
    reencode_spectrograms(args.checkpoint, config, args.input_dir, args.output_dir, args.device, args.batch_size,
                          args.gpus, args.sort_by_length)


if __name__ == '__main__':
    main()
