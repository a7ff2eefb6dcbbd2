"""world_size-2 gloo run of the sharded re-encode host path on CPU (N > 1 coverage)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from mqgan_b200 import reencode as R
rank, world, _ = R.dist_env()
dist.init_process_group("gloo")
def run(batch, lengths):
    return batch * 2.0 + float(batch.shape[1])
done, failed = R.reencode_tree(run, {src!r}, {dst!r}, 3, rank, world, progress=False)
tot, totf = R.finish_distributed(done, failed)
if rank == 0:
    print("TOTAL", tot, totf, flush=True)
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_gloo_reencode(tmp_path):
    from tests.test_host_logic import _make_tree, _fake_run
    from mqgan_b200 import reencode as R
    src, dst, ref = str(tmp_path / "in"), str(tmp_path / "out"), str(tmp_path / "ref")
    _make_tree(src, 17, seed=3)
    R.reencode_tree(_fake_run, src, ref, 3, progress=False)
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, src=src, dst=dst))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "TOTAL 17 0" in res.stdout
    for p in R.list_npy_files(ref):
        assert np.array_equal(np.load(p), np.load(os.path.join(dst, os.path.relpath(p, ref))))
