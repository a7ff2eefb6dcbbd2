#!/usr/bin/env python
"""Drop-in for the reference's reencode_spectrograms.py (same flags, same .npy tree layout;
reference: reencode_spectrograms.py:90-109).  Runs on the sm_100a kernels; ``--gpus N`` is an
additive option that shards the reference's batches over N GPUs of one box (no communication).
Under torchrun (RANK/WORLD_SIZE set) each rank processes its share on its own GPU."""
import argparse
import functools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch  # noqa: E402

from mqgan_b200 import reencode as R  # noqa: E402
from mqgan_b200.scripted_preencoder import ScriptedPreEncoder  # noqa: E402


def _make_model(model_path, device):
    return ScriptedPreEncoder(model_path, device=device)


def reencode_spectrograms(model_path, input_dir, output_dir, device, batch_size, gpus=1, sort_by_length=False):
    print(f"Loading model from: {model_path}")
    rank, world, local = R.dist_env()
    if gpus > 1 and world == 1:
        done, failed = R.run_multi_gpu(functools.partial(_make_model, model_path), input_dir, output_dir,
                                       batch_size, gpus, sort_by_length)
    else:
        if world > 1 and torch.device(device).type == "cuda":
            device = f"cuda:{local}"
            torch.cuda.set_device(local)
        try:
            model = ScriptedPreEncoder(model_path, device=device)
        except (FileNotFoundError, RuntimeError) as e:
            print(f"Error: Could not load the model. {e}")
            return
        print(f"Searching for .npy files in: {input_dir}")

        def run(batch, lengths):
            idx = model.encode(batch, lengths=lengths)
            return model.decode(idx, lengths=lengths)

        done, failed = R.reencode_tree(run, input_dir, output_dir, batch_size, rank, world,
                                        sort_by_length=sort_by_length)
        done, failed = R.finish_distributed(done, failed)
    if rank == 0:
        print("\nProcessing complete.")
        print(f"Re-encoded {done} spectrograms ({failed} failed batches); saved to: {output_dir}")


def main():
    parser = argparse.ArgumentParser(
        description="Re-encode spectrograms using a TorchScript PreEncoder model directory "
                    "(B200-native kernels).", formatter_class=argparse.RawTextHelpFormatter)
    parser.add_argument('--model', type=str, required=True,
                        help='Path to the TorchScript-exported PreEncoder model folder')
    parser.add_argument('--input_dir', type=str, required=True,
                        help='Path to the input folder containing .npy spectrograms.')
    parser.add_argument('--output_dir', type=str, required=True,
                        help='Path to the output folder where re-encoded spectrograms will be saved.')
    parser.add_argument('--device', type=str, default='cpu',
                        help='Device to use for inference (e.g., "cpu", "cuda"). Defaults to \'cpu\' like the '
                             'reference; this build only runs on "cuda" and says so otherwise.')
    parser.add_argument('--batch_size', type=int, default=32,
                        help='Number of spectrograms to process in a single batch. Defaults to 32.')
    parser.add_argument('--sort_by_length', action='store_true',
                        help='(added) batch files of similar length together (less padding, faster). Changes the '
                             'batch composition, which the encoder is sensitive to through padding: off by default.')
    parser.add_argument('--gpus', type=int, default=1,
                        help='(added) shard batches over this many GPUs of one box. Defaults to 1.')
    args = parser.parse_args()
    reencode_spectrograms(args.model, args.input_dir, args.output_dir, args.device, args.batch_size, args.gpus, args.sort_by_length)


if __name__ == '__main__':
    main()
