"""CPU tests of the host side: state-dict surface, loaders, CLI batching/sharding,
C-ABI symbol table (no compute calls - there is no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from mqgan_b200 import _lib, spec as S, reencode as R, ops
from mqgan_b200.preencoder import PreEncoder, get_pre_encoder, sequence_mask
from mqgan_b200.synth import synth_state_dict
from mqgan_b200.engine import folded_weights
from oracle import preencoder_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    _lib.build()
    header = open(os.path.join(ROOT, "include", "mqgan_b200.h")).read()
    declared = set(re.findall(r"\b(mq_[a-z0-9_]+)\s*\(", header))
    declared -= {"mq_stream_t"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.mq_version() == 100
    assert lib.mq_cam_chunks(1000) == 16            # pure host helper, safe without a GPU
    assert isinstance(lib.mq_last_error(), bytes)


def test_ctypes_structs_match_header_sizes(tmp_path):
    """Compile a tiny C program against the header and compare sizeof() with the ctypes mirrors."""
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "mqgan_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n",'
                   'sizeof(mq_conv_params),sizeof(mq_cb2d_params),sizeof(mq_cbam_apply_params),sizeof(mq_fsq_params));return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = list(map(int, subprocess.check_output([str(exe)]).split()))
    assert sizes == [ctypes.sizeof(_lib.ConvParams), ctypes.sizeof(_lib.Cb2dParams),
                     ctypes.sizeof(_lib.CbamApplyParams), ctypes.sizeof(_lib.FsqParams)]


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.MqError, match="no CPU/PyTorch fallback"):
        _lib.lib()


def test_state_dict_surface_and_attributes():
    cfg = S.HIFISPEECH
    m = PreEncoder(cfg.mel_channels, list(cfg.channels), list(cfg.kernel_sizes), fsq_levels=list(cfg.fsq_levels),
                   refiner_base_channels=cfg.refiner_base_channels)
    keys = [k for k, _ in S.param_spec(cfg)]
    assert list(m.state_dict().keys()) == keys and len(keys) == 145
    assert m.codebook_size == 1000 and m.bos_token_id == 1001 and m.eos_token_id == 1002
    assert m.quantizer_dim == 4 and m.refiner_hidden_channels == 16
    assert sum(p.numel() for p in m.parameters()) == 30599524      # 30.60 M
    with pytest.raises(RuntimeError, match="CUDA"):
        m.encode(torch.zeros(1, 8, cfg.mel_channels))
    assert S.flops_per_frame(cfg)["total"] == pytest.approx(738.78e6, rel=1e-4)
    assert S.flops_per_frame(S.HIFIMUSIC)["total"] == pytest.approx(1996.02e6, rel=1e-4)


def test_get_pre_encoder_errors_and_loading(tmp_path):
    cfg = S.TINY
    kw = dict(channels=list(cfg.channels), kernel_sizes=list(cfg.kernel_sizes), mel_channels=cfg.mel_channels,
              fsq_levels=list(cfg.fsq_levels), refiner_base_channels=cfg.refiner_base_channels)
    with pytest.raises(FileNotFoundError):
        get_pre_encoder(str(tmp_path / "missing.pth"), "cpu", **kw)
    sd = synth_state_dict(cfg)
    p = tmp_path / "a.pth"
    torch.save({"not_weights": 1}, p)
    with pytest.raises(KeyError):
        get_pre_encoder(str(p), "cpu", **kw)
    torch.save({"model_state_dict": {"module." + k: v for k, v in sd.items()}}, p)
    m = get_pre_encoder(str(p), "cpu", inference=True, **kw)
    assert not m.training and all(torch.equal(m.state_dict()[k], v) for k, v in sd.items())
    bad = dict(sd)
    bad.pop("proj.bias")
    torch.save({"model_state_dict": bad}, p)
    with pytest.raises(RuntimeError):
        get_pre_encoder(str(p), "cpu", **kw)
    # a state-dict with legacy weight-norm already stripped (inference=True export, App. B4) still loads
    w = folded_weights(sd)
    stripped = {}
    for k, v in sd.items():
        if k.endswith("weight_g"):
            continue
        if k.endswith("weight_v"):
            stripped[k[:-2]] = w[k[:-2]]
        else:
            stripped[k] = v
    torch.save({"model_state_dict": stripped}, p)
    m2 = get_pre_encoder(str(p), "cpu", **kw)
    w2 = folded_weights(m2.state_dict())
    for k in w:
        assert torch.allclose(w[k], w2[k], atol=1e-6), k


def test_folded_weights_equal_oracle_folding():
    sd = synth_state_dict(S.TINY)
    a, b = folded_weights(sd), O.effective_weights(sd)
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)


def ps_check(w):
    """segment-major split packing: segment 5 (x0*w0) holds the leading bf16 term of every tap."""
    ps = ops.pack_conv(w, None, "same1d", split=True)
    K1 = ps.taps * ps.kchunks * 64
    lead = ps.wpack[: w.shape[0], 5 * K1:6 * K1].reshape(w.shape[0], ps.taps, -1)[:, :, : w.shape[1]].permute(0, 2, 1).float()
    rest = ps.wpack[: w.shape[0], 4 * K1:5 * K1].reshape(w.shape[0], ps.taps, -1)[:, :, : w.shape[1]].permute(0, 2, 1).float()
    last = ps.wpack[: w.shape[0], 2 * K1:3 * K1].reshape(w.shape[0], ps.taps, -1)[:, :, : w.shape[1]].permute(0, 2, 1).float()
    return lead + rest + last


def test_weight_packing_layout():
    w = torch.arange(2 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3) / 16
    pc = ops.pack_conv(w, torch.zeros(2), "causal1d", split=False)
    assert (pc.taps, pc.kchunks, pc.nseg, pc.bn, pc.cout_pad) == (3, 1, 1, 32, 32)
    assert pc.tap_dh == [-2, -1, 0] and pc.wpack.shape == (32, 3 * 64)
    assert torch.equal(pc.wpack[1, 64:67].float(), w[1, :, 1])
    assert torch.equal(ps_check(w), w)
    ps = ops.pack_conv(w, None, "same1d", split=True)
    assert ps.tap_dh == [-1, 0, 1] and ps.nseg == 6 and ps.a_coff == [6, 3, 0, 3, 0, 0]
    w0, w1, w2 = ops.split3_bf16(w)
    assert torch.equal(w0.float() + w1.float() + w2.float(), w)
    p2 = ops.pack_conv(torch.randn(768, 64, 3, 3), None, "conv2d3", False)
    assert (p2.bn, p2.cout_pad, p2.taps) == (256, 768, 9) and p2.tap_dh[0] == -1 and p2.tap_dw[2] == 1
    assert ops.choose_bn(384) == (192, 384) and ops.choose_bn(96) == (96, 96) and ops.choose_bn(16) == (32, 32)
    assert ops.choose_tile(1024, 1) == (128, 1) and ops.choose_tile(128, 144) == (8, 16)
    # CTA-pair sub-tiles per CTA on the refiner's shapes (32 utterances, F = 144): the variants the sustained
    # per-launch energy table picks (profiles/conv_bench_r02_sustained.log)
    assert ops.choose_msub_pair(64, 32, 1024, 144, False) == 4        # pre.conv2, up2.conv2
    assert ops.choose_msub_pair(128, 32, 512, 144, False) == 2        # down0.conv1/2, up1.conv2
    assert ops.choose_msub_pair(256, 32, 256, 144, False) == 1        # wide layers: two TMEM buffers
    assert ops.choose_msub_pair(64, 32, 512, 144, True) == 2          # fused up-concat: two skip-parity boxes per slot
    assert ops.choose_msub_pair(32, 32, 1024, 144, False) == 4        # refiner.post as a one-channel 3x3 conv
    assert ops.choose_msub_pair(64, 1, 16, 24, False) == 1            # an image smaller than one pair tile
    # refiner.post packed as the 3x3 convolution it is: one live output channel in an N tile of 32
    pp = ops.pack_conv(torch.randn(1, 64, 3, 3), None, "conv2d3", False)
    assert (pp.bn, pp.cout, pp.cout_pad, pp.taps, pp.kchunks) == (32, 1, 32, 9, 1)


def test_fsq_params_match_reference_constants():
    f = ops.fsq_params([8, 5, 5, 5])
    assert [f.basis[i] for i in range(4)] == [1, 8, 40, 200]
    assert [f.half_w[i] for i in range(4)] == [4, 2, 2, 2]
    assert f.half_l[0] == pytest.approx(3.5035) and f.half_l[1] == pytest.approx(2.002)
    assert f.shift[0] == pytest.approx(0.143695, abs=1e-6) and f.shift[1] == 0.0


def _make_tree(root, n, seed=0):
    rng = np.random.default_rng(seed)
    paths = []
    for i in range(n):
        d = os.path.join(root, f"spk{i % 3}", "sub" if i % 2 else "")
        os.makedirs(d, exist_ok=True)
        p = os.path.join(d, f"utt{i}.npy")
        np.save(p, rng.standard_normal((int(rng.integers(5, 40)), 8)).astype(np.float64 if i % 4 == 0 else np.float32))
        paths.append(p)
    return paths


def _fake_run(batch, lengths):
    # stands in for encode->decode; depends on the batch composition like the real encoder does
    return batch * 2.0 + float(batch.shape[1])


def test_reencode_tree_matches_reference_batching(tmp_path):
    src, dst = str(tmp_path / "in"), str(tmp_path / "out")
    _make_tree(src, 11)
    files = R.list_npy_files(src)
    assert len(files) == 11
    batches = R.make_batches(files, 4)
    assert [len(b) for b in batches] == [4, 4, 3]
    done, failed = R.reencode_tree(_fake_run, src, dst, 4, progress=False)
    assert (done, failed) == (11, 0)
    for b in batches:
        tmax = max(np.load(p).shape[0] for p in b)
        for p in b:
            x = np.load(p)
            y = np.load(os.path.join(dst, os.path.relpath(p, src)))
            assert y.dtype == np.float32 and y.shape == x.shape
            np.testing.assert_allclose(y, x.astype(np.float32) * 2 + tmax)


def test_sharding_is_disjoint_and_order_preserving(tmp_path):
    src = str(tmp_path / "in")
    _make_tree(src, 23, seed=1)
    ref = str(tmp_path / "ref")
    R.reencode_tree(_fake_run, src, ref, 3, progress=False)
    for world in (2, 3):
        dst = str(tmp_path / f"w{world}")
        total = 0
        seen = set()
        for rank in range(world):
            idxs = R.shard_indices(8, rank, world)
            assert not (seen & set(idxs))
            seen |= set(idxs)
            d, f = R.reencode_tree(_fake_run, src, dst, 3, rank, world, progress=False)
            total += d
            assert f == 0
        assert seen == set(range(8)) and total == 23
        for p in R.list_npy_files(ref):
            q = os.path.join(dst, os.path.relpath(p, ref))
            assert np.array_equal(np.load(p), np.load(q))         # byte-identical to the 1-worker run


def test_failing_batch_is_skipped(tmp_path, capsys):
    src, dst = str(tmp_path / "in"), str(tmp_path / "out")
    _make_tree(src, 6)
    calls = {"n": 0}

    def flaky(batch, lengths):
        calls["n"] += 1
        if calls["n"] == 2:
            raise RuntimeError("boom")
        return batch

    done, failed = R.reencode_tree(flaky, src, dst, 2, progress=False)
    assert (done, failed) == (4, 1)
    assert "Could not process batch" in capsys.readouterr().out


def test_sequence_mask_matches_reference_definition():
    lengths = torch.tensor([3, 0, 5])
    m = sequence_mask(5, lengths)
    assert m.tolist() == [[False, False, False, True, True], [True] * 5, [False] * 5]
    assert torch.equal(m, O.sequence_mask(5, lengths))


def test_convert_spectrograms_host_logic(tmp_path):
    """Host side of the mel front-end CLI (convert_spectrograms.py:73-90, :111-120): config validation,
    task discovery, chunking; and the extractor refuses a CPU device (no fallback)."""
    from mqgan_b200 import convert_spectrograms as cs
    cfg = {"io": {"input_folder": str(tmp_path / "in"), "output_folder": str(tmp_path / "out"),
                  "audio_extensions": [".wav", ".flac"]},
           "spectrogram": {"sampling_rate": 44100, "filter_length": 2048, "hop_length": 512, "win_length": 2048,
                           "n_mel_channels": 128, "mel_fmin": 0.0, "mel_fmax": 22050.0}}
    cs.validate_config(cfg)
    bad = {"io": cfg["io"], "spectrogram": {k: v for k, v in cfg["spectrogram"].items() if k != "hop_length"}}
    with pytest.raises(ValueError, match="hop_length"):
        cs.validate_config(bad)
    with pytest.raises(ValueError, match="io"):
        cs.validate_config({"spectrogram": cfg["spectrogram"]})
    os.makedirs(tmp_path / "in" / "spk", exist_ok=True)
    for n in ("a.wav", "b.FLAC", "c.txt"):
        (tmp_path / "in" / "spk" / n).write_bytes(b"")
    tasks = cs.collect_tasks(cfg)
    assert sorted(os.path.basename(t[0]) for t in tasks) == ["a.wav", "b.FLAC"]
    assert all(t[1] == os.path.join(cfg["io"]["output_folder"], "spk") for t in tasks)
    assert cs.chunkify(list(range(7)), 3) == [[0, 1, 2], [3, 4], [5, 6]]
    with pytest.raises(ValueError, match="CUDA"):
        cs.TorchMelSpectrogramExtractor(cfg["spectrogram"], device="cpu")


def test_sort_by_length_is_optional_and_io_pool_preserves_results(tmp_path):
    """--sort_by_length regroups files (less padding) but every file is still processed exactly once and saved
    under its mirrored path; without the flag the reference's os.walk batch composition is untouched.  The
    threaded loads / saves give byte-identical files."""
    rng = np.random.default_rng(5)
    src, dst1, dst2 = str(tmp_path / "in"), str(tmp_path / "o1"), str(tmp_path / "o2")
    lens = {}
    for i in range(11):
        d = os.path.join(src, f"s{i % 3}")
        os.makedirs(d, exist_ok=True)
        T = int(rng.integers(3, 40))
        lens[f"u{i}.npy"] = T
        np.save(os.path.join(d, f"u{i}.npy"), rng.standard_normal((T, 6)).astype(np.float32))
    seen = []

    def run(batch, lengths):
        seen.append(list(lengths))
        return batch * 2.0

    assert R.reencode_tree(run, src, dst1, 4, progress=False) == (11, 0)
    unsorted_batches = [list(b) for b in seen]
    seen.clear()
    assert R.reencode_tree(run, src, dst2, 4, progress=False, sort_by_length=True) == (11, 0)
    flat = [l for b in seen for l in b]
    assert flat == sorted(flat) and sorted(flat) == sorted(lens.values())
    assert [l for b in unsorted_batches for l in b] != flat            # the default order is the walk order
    for p in R.list_npy_files(src):
        rel = os.path.relpath(p, src)
        a, b = np.load(os.path.join(dst1, rel)), np.load(os.path.join(dst2, rel))
        assert a.shape == (lens[os.path.basename(p)], 6) and np.array_equal(a, b) and np.array_equal(a, np.load(p) * 2.0)


# ----------------------------------------------------------------------------
# native .npy I/O of the CLI (mq_npy_probe / mq_npy_read_f32 / mq_npy_write_f32): host code, no GPU needed
# ----------------------------------------------------------------------------
def _npy_tree(root, dtypes=(np.float32, np.float64, np.float16), n=7, n_mels=12):
    os.makedirs(os.path.join(root, "a", "b"), exist_ok=True)
    g = np.random.default_rng(3)
    paths = []
    for i in range(n):
        d = root if i % 3 == 0 else os.path.join(root, "a") if i % 3 == 1 else os.path.join(root, "a", "b")
        arr = (g.standard_normal((5 + 7 * i, n_mels)) * 3 - 4).astype(dtypes[i % len(dtypes)])
        p = os.path.join(d, f"u{i}.npy")
        np.save(p, arr)
        paths.append(p)
    return paths


def test_native_npy_reader_equals_numpy(tmp_path, monkeypatch):
    from mqgan_b200 import reencode as R
    paths = _npy_tree(str(tmp_path))
    monkeypatch.setattr(R, "NATIVE_IO", True)
    a, la = R.load_and_pad(paths)
    monkeypatch.setattr(R, "NATIVE_IO", False)
    b, lb = R.load_and_pad(paths)
    assert la == lb and a.dtype == torch.float32 and torch.equal(a, b)
    # pooled staging buffer variant: same content, buffer handed back for reuse
    monkeypatch.setattr(R, "NATIVE_IO", True)
    c, lc, handle = R.load_and_pad(paths, pooled=True)
    assert torch.equal(c, b) and handle is not None and handle.numel() >= c.numel()
    R._pinned.put(handle)
    assert any(t is handle for t in R._pinned._free)
    # a file numpy can read but the library does not handle (integers, Fortran order) -> whole batch through numpy
    np.save(os.path.join(str(tmp_path), "ints.npy"), np.arange(24).reshape(2, 12))
    np.save(os.path.join(str(tmp_path), "fort.npy"), np.asfortranarray(np.ones((3, 12), np.float32)))
    mixed = paths[:2] + [os.path.join(str(tmp_path), "ints.npy"), os.path.join(str(tmp_path), "fort.npy")]
    d, ld = R.load_and_pad(mixed)
    assert ld[-2:] == [2, 3] and float(d[2, 1, 11]) == 23.0 and float(d[3, :3].min()) == 1.0
    # mismatching mel width is an error either way
    np.save(os.path.join(str(tmp_path), "wide.npy"), np.zeros((4, 13), np.float32))
    with pytest.raises(ValueError):
        R.load_and_pad(paths[:1] + [os.path.join(str(tmp_path), "wide.npy")])
    with pytest.raises(OSError):
        R.load_and_pad([os.path.join(str(tmp_path), "missing.npy")])


def test_native_npy_writer_is_byte_identical_to_numpy(tmp_path, monkeypatch):
    from mqgan_b200 import reencode as R
    inp, out_a, out_b = str(tmp_path / "in"), str(tmp_path / "native"), str(tmp_path / "numpy")
    paths = _npy_tree(inp, dtypes=(np.float32,))
    batch, lengths = R.load_and_pad(paths)
    monkeypatch.setattr(R, "NATIVE_IO", True)
    R.save_outputs(batch, lengths, paths, inp, out_a)
    monkeypatch.setattr(R, "NATIVE_IO", False)
    R.save_outputs(batch, lengths, paths, inp, out_b)
    for p in paths:
        rel = os.path.relpath(p, inp)
        a = open(os.path.join(out_a, rel), "rb").read()
        assert a == open(os.path.join(out_b, rel), "rb").read()
        assert a == open(p, "rb").read()                       # float32 in -> identical file out (identity "model")
        assert np.load(os.path.join(out_a, rel)).dtype == np.float32


def test_reencode_tree_identity_model_round_trips_files(tmp_path):
    from mqgan_b200 import reencode as R
    inp, out = str(tmp_path / "in"), str(tmp_path / "out")
    paths = _npy_tree(inp, n=11)
    done, failed = R.reencode_tree(lambda batch, lengths: batch.clone(), inp, out, batch_size=4, progress=False)
    assert (done, failed) == (11, 0)
    for p in paths:
        got = np.load(os.path.join(out, os.path.relpath(p, inp)))
        assert got.dtype == np.float32 and np.array_equal(got, np.load(p).astype(np.float32))


# ----------------------------------------------------------------------------
# training-step configuration (spec.py) against the reference's YAML and parameter counts
# ----------------------------------------------------------------------------
def test_discriminator_configs_match_reference_yaml_and_sizes():
    from mqgan_b200.synth import synth_disc_state_dict
    patch_yaml = {"hidden_channels": [256, 256, 384, 512, 512], "kernel_sizes": [5, 5, 5, 3, 3, 3],
                  "strides": [[1, 2], [2, 2], [2, 2], [2, 1], [2, 1], [2, 1]]}          # configs/model_config_hifispeech.yaml:23-26
    mb_yaml = {"hidden_channels": [128, 128, 256, 256, 384], "kernel_sizes": [7, 5, 3, 3, 3, 3], "n_bins": 8, "n_no_strides": 2}
    assert S.PatchDiscConfig.from_patch_yaml(128, patch_yaml) == S.HIFISPEECH_PATCH_D
    assert S.MultiBinConfig.from_yaml(128, mb_yaml) == S.HIFISPEECH_MULTIBIN_D
    # the logits conv always runs at stride 1 (discriminators.py:160-170); bands use (3, k) kernels and time-only strides
    assert S.HIFISPEECH_PATCH_D.layer_stride(5) == (1, 1) and S.HIFISPEECH_PATCH_D.layer_stride(1) == (2, 2)
    bc = S.HIFISPEECH_MULTIBIN_D.bin_config
    assert bc.mel_channels == 16 and bc.kernels[0] == (3, 7) and bc.strides[:3] == ((1, 1), (1, 1), (1, 2))
    assert S.HIFISPEECH_PATCH_D.feature_layers == (False, False, True, True, True, False)       # discriminators.py:108-112
    n_pd = sum(v.numel() for k, v in synth_disc_state_dict(S.patch_disc_param_spec(S.HIFISPEECH_PATCH_D)).items()
               if not S.is_disc_buffer(k))
    n_mb = sum(v.numel() for k, v in synth_disc_state_dict(S.multibin_param_spec(S.HIFISPEECH_MULTIBIN_D)).items()
               if not S.is_disc_buffer(k))
    assert abs((n_pd + n_mb) / 1e6 - 24.8) < 0.1                  # 24.8 M discriminator parameters
    with pytest.raises(ValueError):
        S.MultiBinConfig(48, 3, (16, 24), (3, 3, 3), 2)           # hidden size must divide n_bins (discriminators.py:264)
    with pytest.raises(ValueError):
        S.PatchDiscConfig(32, (16,), ((3, 3),), ((1, 1),))        # kernel_sizes must be hidden_channels + 1


def test_dgrad_weight_and_taps_are_consistent_on_cpu():
    """ops.dgrad_weight / ops.conv_taps (host logic of the training step): the mirrored, transposed weight applied as a
    plain correlation equals autograd's data gradient (float64, CPU)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 6, 9, 5, generator=g, dtype=torch.float64, requires_grad=True)          # (N, Cin, H, W)
    w = torch.randn(4, 6, 3, 3, generator=g, dtype=torch.float64)
    y = F.conv2d(x, w, padding=1)
    dy = torch.randn_like(y)
    y.backward(dy)
    wd, kind = ops.dgrad_weight(w, "conv2d3")
    assert kind == "conv2d3" and wd.shape == (6, 4, 3, 3)
    assert torch.allclose(F.conv2d(dy, wd, padding=1), x.grad, atol=1e-12)
    x1 = torch.randn(2, 6, 11, generator=g, dtype=torch.float64, requires_grad=True)
    w1 = torch.randn(4, 6, 5, generator=g, dtype=torch.float64)
    y1 = F.conv1d(F.pad(x1, (4, 0)), w1)                                                        # causal (attentions.py:471-474)
    dy1 = torch.randn_like(y1)
    y1.backward(dy1)
    wd1, kind1 = ops.dgrad_weight(w1, "causal1d")
    assert kind1 == "anticausal1d"
    assert torch.allclose(F.conv1d(F.pad(dy1, (0, 4)), wd1), x1.grad, atol=1e-12)               # taps at rows 0 .. k-1
    assert ops.conv_taps("causal1d", w1.shape) == ([-4, -3, -2, -1, 0], [0] * 5)
    assert ops.conv_taps("conv2d3", w.shape)[0] == [-1, -1, -1, 0, 0, 0, 1, 1, 1]


def test_length_groups_partition_rules():
    """PreEncoderEngine._length_groups (host logic of decode(lengths=...)): a partition of the batch into length-sorted
    groups whose T is the longest member rounded up to 8 frames plus one coarse row, capped by the batch T and by
    max_chunk_frames; None when nothing is saved."""
    import types
    from mqgan_b200.engine import PreEncoderEngine
    stub = types.SimpleNamespace(max_chunk_frames=32768, group_cost_frames=2048, cfg=types.SimpleNamespace(refiner_depth=3))
    groups = PreEncoderEngine._length_groups
    rng = np.random.default_rng(0)
    lens = rng.integers(300, 1100, size=32).tolist()
    T = max(lens)
    g = groups(stub, lens, T)
    assert g is not None and len(g) >= 2
    assert sorted(i for m, _ in g for i in m) == list(range(32))
    for members, Tg in g:
        longest = max(lens[i] for i in members)
        assert Tg % 8 == 0 or Tg == T
        assert Tg == T or Tg >= (-(-longest // 8) + 1) * 8          # a whole coarse row of padding follows the longest
        assert Tg <= T and len(members) * Tg <= 32768
    computed = sum(len(m) * Tg for m, Tg in g)
    assert computed < 0.9 * 32 * T                                  # the point of it: fewer padded frames
    assert groups(stub, [T] * 32, T) is None                        # rectangular batch
    assert groups(stub, [500], 500) is None
    assert groups(stub, [1000, 990, 1010, 1005], 1010) is None      # nearly rectangular: not worth another launch group
    stub.max_chunk_frames = 4096                                    # groups are split to fit the refiner chunk size
    g2 = groups(stub, lens, T)
    assert all(len(m) * Tg <= 4096 or len(m) == 1 for m, Tg in g2)


def test_native_npy_io_property_random_shapes_and_dtypes(tmp_path):
    """Property test (hypothesis) of mq_npy_read_f32 / mq_npy_write_f32 against numpy over random shapes, dtypes, header
    versions and padding targets."""
    import ctypes as C
    from hypothesis import given, settings, strategies as st
    import numpy.lib.format as nf
    lib = _lib.lib()
    path = str(tmp_path / "p.npy").encode()

    @settings(max_examples=60, deadline=None)
    @given(rows=st.integers(0, 40), cols=st.integers(1, 33), dt=st.sampled_from(["<f4", "<f8", "<f2"]),
           version=st.sampled_from([(1, 0), (2, 0), (3, 0)]), extra=st.integers(0, 5), seed=st.integers(0, 2 ** 16))
    def check(rows, cols, dt, version, extra, seed):
        arr = (np.random.default_rng(seed).standard_normal((rows, cols)) * 100).astype(np.dtype(dt))
        with open(path, "wb") as f:
            nf.write_array(f, arr, version=version)
        r, c, d = C.c_int64(), C.c_int64(), C.c_int()
        assert lib.mq_npy_probe(path, C.byref(r), C.byref(c), C.byref(d), None) == 0
        assert (r.value, c.value, d.value) == (rows, cols, {"<f4": 0, "<f8": 1, "<f2": 2}[dt])
        dst = np.full((rows + extra, cols), np.nan, np.float32)
        got_rows = C.c_int64()
        assert lib.mq_npy_read_f32(path, dst.ctypes.data, rows + extra, cols, C.byref(got_rows)) == 0
        assert got_rows.value == rows
        assert np.array_equal(dst[:rows], arr.astype(np.float32)) and (dst[rows:] == 0).all()
        if rows > 1:                                   # destination shorter than the file: the head is read
            short = np.full((rows - 1, cols), np.nan, np.float32)
            assert lib.mq_npy_read_f32(path, short.ctypes.data, rows - 1, cols, None) == 0
            assert np.array_equal(short, arr[: rows - 1].astype(np.float32))
        out = str(tmp_path / "w.npy")
        a32 = np.ascontiguousarray(arr.astype(np.float32))
        assert lib.mq_npy_write_f32(out.encode(), a32.ctypes.data, rows, cols) == 0
        ref = str(tmp_path / "r.npy")
        np.save(ref, a32)
        assert open(out, "rb").read() == open(ref, "rb").read()

    check()
    bad = np.full((2, 3), np.nan, np.float32)
    assert lib.mq_npy_read_f32(path, bad.ctypes.data, 2, 999, None) in (1, 4, 5)      # column mismatch is reported, not read


def test_refiner_post_row_sum_decomposition():
    """The engine runs refiner.post (C -> 1, 3x3; reference preencoder.py:191) as a one-row convolution with three output
    channels (channel dt + 1 = kernel row dt along F) whose outputs the tail adds down T: P0[t-1] + P1[t] + P2[t+1].
    Same weights, same zero padding - checked here against F.conv2d with the packing the engine uses."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    C, T, Fq = 8, 11, 13
    w = torch.randn(1, C, 3, 3, generator=g, dtype=torch.float64)
    x = torch.randn(2, C, T, Fq, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, w, padding=1)[:, 0]
    wr = torch.zeros(4, C, 3, dtype=torch.float64)
    wr[:3] = w[0].permute(1, 0, 2)                                   # engine.py: (dt, C, df)
    P = F.conv2d(x, wr.unsqueeze(2), padding=(0, 1))                 # (B, 4, T, F): a 1x3 convolution per output channel
    Pz = F.pad(P, (0, 0, 1, 1))                                      # zero rows above and below the image
    got = Pz[:, 0, 0:T] + Pz[:, 1, 1:T + 1] + Pz[:, 2, 2:T + 2]      # P0[t-1] + P1[t] + P2[t+1]
    assert torch.allclose(got, ref, atol=1e-12, rtol=0)
    assert float(P[:, 3].abs().max()) == 0.0                         # the fourth (padding) channel stays zero
    pc = ops.pack_conv(wr.float(), None, "taps2d", False, taps=([0, 0, 0], [-1, 0, 1]))
    assert (pc.taps, pc.tap_dh[:3], pc.tap_dw[:3], pc.cout, pc.bn) == (3, [0, 0, 0], [-1, 0, 1], 4, 32)
    # K order of the packed weight: tap-major, then channel (padded to 64)
    assert torch.equal(pc.wpack[2, 64:64 + C].float(), wr[2, :, 1].float().to(torch.bfloat16).float())
