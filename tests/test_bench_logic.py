"""Host-side logic of bench.py and of the checker / reference plumbing (no GPU, no compute through the library)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import gpu_checker as G  # noqa: E402
from oracle import preencoder_oracle as O  # noqa: E402
from oracle import reference_runner as RR  # noqa: E402


def test_both_arms_describe_the_same_config():
    """`config` of the JSON line is the same object in the CUDA arm and in --impl reference (the driver compares them)."""
    for name in bench.WORKLOADS:
        a, b = bench.config_dict(name, 1), bench.config_dict(name, 1)
        assert a == b and json.loads(json.dumps(a)) == a
        assert a["workload"].startswith(name) and a["batch_per_gpu"] == bench.WORKLOADS[name][1]
    head = bench.config_dict(bench.DEFAULT_WORKLOAD, 8)
    assert "bf16 decoder" in head["workload"] and "x8" in head["parallelism"]
    assert bench.WORKLOADS["hifispeech_256x1024_fp32"][4] == "f16x2"          # the fp32-grade decoder line
    assert bench.CPU_CONFIG0 == (16, 512)                                       # BASELINE configs[0]


def test_reference_install_and_lookup(tmp_path, monkeypatch):
    src = tmp_path / "ref"
    (src / "configs").mkdir(parents=True)
    (src / "preencoder.py").write_text("X = 1\n")
    (src / "quantizer.py").write_text("Y = 2\n")
    (src / "configs" / "model.yaml").write_text("a: 1\n")
    (src / "notes.txt").write_text("not copied\n")
    dst = tmp_path / "baseline" / "_ref"
    assert RR.install(str(src), str(dst)) == str(dst)
    assert sorted(os.listdir(dst)) == ["configs", "preencoder.py", "quantizer.py"]
    assert RR.install(str(tmp_path / "missing"), str(dst)) is None               # nothing to install from
    monkeypatch.setenv("MQGAN_REFERENCE", str(dst))
    assert RR.find() == str(dst)


def test_gitignore_keeps_the_reference_copy_out_of_history():
    lines = open(os.path.join(ROOT, ".gitignore")).read().split()
    assert "baseline/_ref/" in lines
    gi = os.path.join(ROOT, ".gpurunignore")
    assert not os.path.exists(gi) or "baseline" not in open(gi).read()            # it has to travel to the GPU box


@pytest.mark.parametrize("levels", [[8, 5, 5, 5], [8, 8, 5, 5, 5]])
def test_checker_quantiser_equals_reference_on_adversarial_rows(levels):
    fx = np.load(os.path.join(ROOT, "tests", "golden", "fsq_" + "_".join(map(str, levels)) + ".npz"))
    z = torch.from_numpy(fx["z_adv"])
    idx, margin = G.fsq_indices_and_margin(z, levels)
    assert torch.equal(idx, torch.from_numpy(fx["indices_adv"].astype(np.int64)))   # same fp32 arithmetic as the reference
    assert float(margin.min()) < 1e-3                                                # and the rows really sit on boundaries
    assert torch.equal(idx, O.fsq_quantize(z, levels)[1])


def test_sharded_cli_caps_host_threads(monkeypatch):
    """Eight workers must not bring up eight machine-wide intra-op pools (the limiter measured at 8 GPUs)."""
    from mqgan_b200 import reencode as R
    monkeypatch.delenv("MQ_WORKER_THREADS", raising=False)
    monkeypatch.delenv("MQ_IO_THREADS", raising=False)
    before_threads, before_io, before_pool = torch.get_num_threads(), R.IO_THREADS, R._io_pool
    try:
        R._io_pool = None
        R._cap_host_threads(8)
        assert torch.get_num_threads() <= 4 and R.IO_THREADS == 1
        R.IO_THREADS = 4
        R._cap_host_threads(1)
        assert torch.get_num_threads() <= 4 and R.IO_THREADS == 4
        monkeypatch.setenv("MQ_WORKER_THREADS", "0")
        torch.set_num_threads(before_threads)
        R._cap_host_threads(8)
        assert torch.get_num_threads() == before_threads                             # 0 = leave torch's default
    finally:
        torch.set_num_threads(before_threads)
        R.IO_THREADS, R._io_pool = before_io, before_pool


def test_amplified_weights_are_a_pure_function_of_the_seed():
    from mqgan_b200 import spec as S
    from mqgan_b200.synth import amplify_state_dict, synth_state_dict
    sd = synth_state_dict(S.TINY, 0)
    a, b = amplify_state_dict(sd), amplify_state_dict(synth_state_dict(S.TINY, 0))
    assert all(torch.equal(a[k], b[k]) for k in a) and set(a) == set(sd)
    g = "encoder_blocks.0.conv1.parametrizations.weight.original0"
    assert torch.allclose(a[g], sd[g] * 2.2) and torch.allclose(a["proj.weight"], sd["proj.weight"] * 8.0)
    assert torch.allclose(a["out_proj.bias"], sd["out_proj.bias"] - 4.0)
    r = "refiner.mid.conv1.parametrizations.weight.original0"
    assert torch.allclose(a[r], sd[r] * 2.35)
