"""Writes a tiny synthetic dataset + config for a smoke run of the drop-in train.py (used to check the data-parallel
driver under torchrun).  Usage: python tools/train_cli_smoke.py <dir>  -> prints the config path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yaml

from tests.test_train_cli import _tiny_config, _write_mels

root = sys.argv[1]
data, out = os.path.join(root, "mels"), os.path.join(root, "run")
_write_mels(data, 70, 32, [40, 60, 25, 90, 55])
cfg = _tiny_config(data, out)
cfg["training"]["num_epochs"] = 3
cfg["training"]["discriminator_train_start_epoch"] = 1
path = os.path.join(root, "cfg.yaml")
with open(path, "w") as f:
    yaml.safe_dump(cfg, f)
print(path)
