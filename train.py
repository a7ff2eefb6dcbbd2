#!/usr/bin/env python
"""Drop-in for the reference's ``train.py`` (same flags and YAML; see mqgan_b200/train_cli.py).

    python train.py --config configs/model_config_hifispeech.yaml [--pretrained ckpt.pth] [--output_dir logs/run]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 train.py --config ...
"""
import sys

from mqgan_b200.train_cli import main

if __name__ == "__main__":
    sys.exit(main())
