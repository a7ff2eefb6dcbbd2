"""End-to-end parity of the CUDA path (PreEncoder API -> C ABI) against the CPU
oracle and the committed reference outputs (tests/golden/).

Gates (BASELINE north_star):
  * fp32-grade modes ("f16x2" = two fp16 terms per operand, the default; "bf16x3" = three bf16
    terms): indices equal to the reference on every frame whose float64 distance to an FSQ
    rounding boundary exceeds TAU; raw agreement reported and required >= 99.5 %.
  * bf16 encoder mode: agreement rate reported, required >= 80 %.
  * reconstructed mels, default decoder (bf16 operands, fp32 accumulate, bf16 activations):
    max-abs error <= MEL_ATOL + MEL_RTOL * max|ref| and relative L2 <= MEL_REL_L2 - about 3x the measured error
    (1.0-1.2e-3 max-abs at |ref| <= 0.5, rel-L2 2e-3), also on the amplified fixtures (|ref| up to 20);
  * reconstructed mels, fp32-grade decoder (decoder_precision="f16x2"): max-abs <= MEL32_RTOL * max(1, max|ref|),
    relative L2 <= MEL32_REL_L2 (1e-4, the reference's own fp32 noise level).
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from mqgan_b200.preencoder import PreEncoder, sequence_mask  # noqa: E402
from mqgan_b200.synth import synth_mels, synth_lengths  # noqa: E402
from oracle import preencoder_oracle as O  # noqa: E402
from tests.helpers import load_golden, index_report  # noqa: E402

TAU = 2e-4          # bounded-latent units (rounding boundaries are 1 apart)
Z_ATOL = 5e-5       # pre-quantiser latents (unit std) vs float64; measured 4e-6 .. 1.4e-5 (reference fp32: 4e-6 .. 9e-6)
MEL_ATOL = 3e-3
MEL_RTOL = 3e-3
MEL_RTOL_AMP = 1.5e-2   # amplified fixtures (|mel| up to 20, APTx saturated): measured 4.6e-3 .. 6.1e-3 of max|ref| with bf16
Z_ATOL_AMP = 2e-4       # amplified fixtures: activations 10-100x larger; measured 7.6e-5 (indices: 0 mismatches)
MEL_REL_L2 = 5e-3
MEL32_RTOL = 1e-4
MEL32_REL_L2 = 1e-4

REPORT = {}


def _dump():
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_report.json"), "w") as f:
        json.dump(REPORT, f, indent=1)


def _model(cfg, sd, precision="f16x2", decoder_precision="bf16"):
    m = PreEncoder(cfg.mel_channels, list(cfg.channels), list(cfg.kernel_sizes), fsq_levels=list(cfg.fsq_levels),
                   dropout=0.0, refiner_base_channels=cfg.refiner_base_channels, refiner_depth=cfg.refiner_depth,
                   refiner_hidden_proj_divisor=cfg.refiner_hidden_proj_divisor, encoder_precision=precision,
                   decoder_precision=decoder_precision)
    m.load_state_dict(sd, strict=True)
    return m.to("cuda").eval()


@pytest.mark.parametrize("precision", ["f16x2", "bf16x3"])
@pytest.mark.parametrize("name", ["tiny", "hifispeech", "hifimusic", "tiny_amp", "hifispeech_amp"])
def test_encode_indices_vs_reference(name, precision):
    cfg, sd, mel, lengths, fx = load_golden(name)
    T = mel.shape[1]
    mask = sequence_mask(T, lengths).unsqueeze(1)
    model = _model(cfg, sd, precision)
    idx, z = model.engine().encode(mel.cuda(), mask.cuda(), return_latents=True)
    ref_idx = torch.from_numpy(fx["indices"].astype(np.int64))
    ref_z = torch.from_numpy(fx["z"])
    z64 = O.encode_latents(sd, cfg, mel, mask, dtype=torch.float64)
    margin = O.fsq_round_margin(z64, cfg.fsq_levels)
    rep = index_report(idx, ref_idx, margin, TAU)
    rep["z_maxabs_vs_ref32"] = float((z.cpu() - ref_z).abs().max())
    rep["z_maxabs_vs_fp64"] = float((z.cpu().double() - z64).abs().max())
    rep["ref32_maxabs_vs_fp64"] = float((ref_z.double() - z64).abs().max())
    REPORT[f"encode/{name}/{precision}"] = rep
    _dump()
    print(name, precision, rep)
    assert rep["safe_mismatch"] == 0, rep
    assert rep["agree"] >= 0.995, rep
    # latents (unit std) within Z_ATOL of float64 (Z_ATOL_AMP on the amplified fixtures, whose activations are
    # 10-100x larger); the reference's own fp32 error is reported beside it
    assert rep["z_maxabs_vs_fp64"] <= (Z_ATOL_AMP if name.endswith("_amp") else Z_ATOL), rep
    # API contract: (B, T) int64 on the module's device
    out = model.encode(mel, mask)
    assert out.dtype == torch.int64 and tuple(out.shape) == tuple(mel.shape[:2]) and out.is_cuda
    assert torch.equal(out, idx)


@pytest.mark.parametrize("decoder_precision", ["bf16", "f16x2"])
@pytest.mark.parametrize("name", ["tiny", "hifispeech", "hifimusic", "tiny_amp", "hifispeech_amp"])
def test_decode_mels_vs_reference(name, decoder_precision):
    cfg, sd, mel, lengths, fx = load_golden(name)
    T = mel.shape[1]
    mask = sequence_mask(T, lengths).unsqueeze(1)
    model = _model(cfg, sd, decoder_precision=decoder_precision)
    for ik, rk in (("indices", "recon"), ("rand_indices", "rand_recon")):
        idx = torch.from_numpy(fx[ik].astype(np.int64))
        ref = torch.from_numpy(fx[rk])
        out = model.decode(idx.cuda(), mask.cuda()).cpu()
        assert out.shape == ref.shape and out.dtype == torch.float32
        err = float((out - ref).abs().max())
        scale = float(ref.abs().max())
        rel = float((out - ref).norm() / ref.norm())
        REPORT[f"decode/{name}/{ik}/{decoder_precision}"] = {"max_abs_err": err, "ref_max": scale, "rel_l2": rel}
        _dump()
        print(name, ik, decoder_precision, REPORT[f"decode/{name}/{ik}/{decoder_precision}"])
        if decoder_precision == "bf16":
            assert err <= MEL_ATOL + (MEL_RTOL_AMP if name.endswith("_amp") else MEL_RTOL) * scale, (err, scale)
            assert rel <= MEL_REL_L2, rel
        else:
            assert err <= MEL32_RTOL * max(1.0, scale), (err, scale)
            assert rel <= MEL32_REL_L2, rel
    # return_hidden: (x_post, last_hid (B, C0, T))
    x_post, hid = model.decode(idx.cuda(), mask.cuda(), return_hidden=True)
    assert tuple(hid.shape) == (mel.shape[0], cfg.c0, T)
    # padded frames are not zeroed (App. B8)
    pad = mask.squeeze(1)
    if pad.any():
        assert float(x_post.cpu()[pad].abs().max()) > 0


def test_no_mask_and_forward():
    cfg, sd, mel, lengths, fx = load_golden("tiny")
    model = _model(cfg, sd)
    idx = model.encode(mel[:1].cuda(), None)
    ref = torch.from_numpy(fx["nomask_indices"].astype(np.int64))
    assert float((idx.cpu() == ref).float().mean()) >= 0.99
    out = model.decode(idx, None).cpu()
    refm = torch.from_numpy(fx["nomask_recon"])
    if torch.equal(idx.cpu(), ref):
        assert float((out - refm).abs().max()) <= MEL_ATOL + MEL_RTOL * float(refm.abs().max())
    x_recon, x_post = model(mel.cuda(), lengths.cuda())
    assert x_recon.shape == x_post.shape == mel.shape


def test_bf16_encoder_mode_reports_agreement():
    cfg, sd, mel, lengths, fx = load_golden("hifispeech")
    mask = sequence_mask(mel.shape[1], lengths).unsqueeze(1)
    model = _model(cfg, sd, precision="bf16")
    idx = model.encode(mel.cuda(), mask.cuda())
    rep = index_report(idx, torch.from_numpy(fx["indices"].astype(np.int64)))
    REPORT["encode_bf16/hifispeech"] = rep
    _dump()
    print("bf16 encoder agreement", rep)
    assert rep["agree"] >= 0.80


def test_batch_invariance_and_determinism():
    """Utterances are independent given the padded length: encoding a
    sub-batch alone gives the same indices bit for bit; chunked execution too."""
    cfg, sd, mel, lengths, fx = load_golden("tiny")
    T = mel.shape[1]
    mask = sequence_mask(T, lengths).unsqueeze(1)
    model = _model(cfg, sd)
    a = model.encode(mel.cuda(), mask.cuda())
    b = model.encode(mel[1:3].cuda(), mask[1:3].cuda())
    assert torch.equal(a[1:3], b)
    assert torch.equal(a, model.encode(mel.cuda(), mask.cuda()))
    eng = model.engine()
    old = eng.max_chunk_frames, eng.max_chunk_frames_enc
    eng.max_chunk_frames = eng.max_chunk_frames_enc = T            # one utterance per chunk
    try:
        assert torch.equal(a, model.encode(mel.cuda(), mask.cuda()))
        d1 = model.decode(a, mask.cuda())
    finally:
        eng.max_chunk_frames, eng.max_chunk_frames_enc = old
    d2 = model.decode(a, mask.cuda())
    assert torch.equal(d1, d2)


def test_decode_in_length_groups_is_bit_identical():
    """decode(lengths=host ints): a ragged batch goes through the refiner in length-sorted groups cut to their own
    longest utterance; decoder + refiner are padding-invariant, so every frame - valid and padded -
    equals the plain padded-batch result bit for bit."""
    cfg, sd, _, _, _ = load_golden("tiny")
    B, T = 12, 160
    lengths = torch.tensor([160, 31, 96, 17, 150, 64, 40, 8, 121, 77, 33, 5])
    mask = sequence_mask(T, lengths).unsqueeze(1)
    g = torch.Generator().manual_seed(3)
    idx = torch.randint(0, cfg.codebook_size, (B, T), generator=g).cuda()
    model = _model(cfg, sd)
    eng = model.engine()
    ref = model.decode(idx, mask.cuda())
    eng.group_cost_frames = 0                                      # tiny batch: group whenever it saves frames
    groups = eng._length_groups(lengths.tolist(), T)
    assert groups is not None and len(groups) > 1
    assert sorted(i for members, _ in groups for i in members) == list(range(B))
    assert all(Tg % 8 == 0 and (Tg == T or Tg >= max(int(lengths[i]) for i in members) + 8) for members, Tg in groups)
    got = model.decode(idx, mask.cuda(), lengths=lengths.tolist())
    diff = (got - ref).abs()
    print("length-group decode: max abs diff", float(diff.max()), "frames differing", int((diff.amax(dim=2) > 0).sum()),
          "of", B * T, "max |ref|", float(ref.abs().max()))
    assert torch.equal(got, ref)
    host = torch.empty(B, T, cfg.mel_channels).pin_memory()
    got2 = model.decode(idx, mask.cuda(), lengths=lengths, host_out=host)
    torch.cuda.synchronize()
    assert torch.equal(got2, ref) and torch.equal(host, ref.cpu())
    assert eng._length_groups([T] * B, T) is None                  # nothing to save on a full-length batch
    with pytest.raises(ValueError):
        model.decode(idx, None, lengths=lengths.tolist())


@pytest.mark.parametrize("decoder_precision", ["bf16", "f16x2"])
def test_refiner_post_direct_conv_matches_tap_planes(decoder_precision):
    """refiner.post (C -> 1, 3x3; reference preencoder.py:191) as three row sums (default) or as one output channel of the
    CTA-pair halo kernel, against the older formulation (a 1x1 GEMM into nine tap planes that the tail shift-adds):
    the same sum in another order."""
    cfg, sd, _, lengths, _ = load_golden("hifispeech")
    B, T = 3, 200
    lens = torch.tensor([200, 133, 57])
    mask = sequence_mask(T, lens).unsqueeze(1).cuda()
    idx = torch.randint(0, cfg.codebook_size, (B, T), generator=torch.Generator().manual_seed(11)).cuda()
    outs = {}
    for mode in ("rows", "direct", "planes"):
        eng = _model(cfg, sd, decoder_precision=decoder_precision).engine()
        eng.post_mode = mode
        outs[mode] = eng.decode(idx, mask)
    ref = outs["planes"]
    for mode in ("rows", "direct"):
        err = float((outs[mode] - ref).abs().max())
        print("post", mode, "vs tap planes: max abs diff", err, "max |ref|", float(ref.abs().max()))
        # bf16 decoder: identical bf16 inputs, fp32 sums in another order; f16x2: three product segments per tap
        assert err <= (2e-5 if decoder_precision == "bf16" else 2e-6) * max(1.0, float(ref.abs().max()))
        assert torch.equal(outs[mode][1, 133:], ref[1, 133:])          # masked frames: x_recon only, every way


def test_decode_in_length_groups_hifispeech_sizes():
    """The same at the real model's widths and CLI-like lengths."""
    cfg, sd, _, _, _ = load_golden("hifispeech")
    B, T = 16, 600
    lengths = torch.tensor([600, 212, 433, 318, 590, 255, 377, 201, 512, 466, 289, 344, 230, 571, 405, 263])
    mask = sequence_mask(T, lengths).unsqueeze(1)
    g = torch.Generator().manual_seed(4)
    idx = torch.randint(0, cfg.codebook_size, (B, T), generator=g).cuda()
    model = _model(cfg, sd)
    model.engine().group_cost_frames = 128          # 16 utterances are too few for the default cost model to split
    groups = model.engine()._length_groups(lengths.tolist(), T)
    assert groups is not None and len(groups) >= 2
    ref = model.decode(idx, mask.cuda())
    got = model.decode(idx, mask.cuda(), lengths=lengths.tolist())
    assert torch.equal(got, ref)


def test_decode_streams_result_to_pinned_host_buffer():
    """decode(..., host_out=pinned) copies every refiner chunk to the host while later chunks compute; the host
    buffer equals the returned device tensor once decode has returned and the stream is synchronised."""
    cfg, sd, mel, lengths, fx = load_golden("tiny")
    T = mel.shape[1]
    mask = sequence_mask(T, lengths).unsqueeze(1)
    model = _model(cfg, sd)
    idx = model.encode(mel.cuda(), mask.cuda())
    eng = model.engine()
    old = eng.max_chunk_frames
    eng.max_chunk_frames = T                                      # one utterance per refiner chunk -> several copies
    try:
        host = torch.full((mel.shape[0], T, cfg.mel_channels), float("nan")).pin_memory()
        dev_out = model.decode(idx, mask.cuda(), host_out=host)
        torch.cuda.synchronize()
        assert torch.equal(host, dev_out.cpu())
        assert torch.equal(dev_out, model.decode(idx, mask.cuda()))
    finally:
        eng.max_chunk_frames = old
    with pytest.raises(ValueError):
        model.decode(idx, mask.cuda(), host_out=torch.empty(mel.shape[0], T, cfg.mel_channels))     # not pinned


def test_hifispeech_longer_ragged_batch_vs_oracle():
    """A larger seeded case than the fixtures (B=4, T=200, ragged), checked against the oracle run here."""
    cfg, sd, _, _, _ = load_golden("hifispeech")
    B, T = 4, 200
    mel = synth_mels(B, T, cfg.mel_channels, seed=7)
    lengths = synth_lengths(B, T, seed=7)
    pad = torch.arange(T)[None, :] >= lengths[:, None]
    mel = mel.masked_fill(pad.unsqueeze(-1), 0.0)
    mask = pad.unsqueeze(1)
    w = O.effective_weights(sd)
    z32 = O.encode_latents(w, cfg, mel, mask, folded=True)
    ref_idx = O.fsq_quantize(z32, cfg.fsq_levels)[1]
    z64 = O.encode_latents(sd, cfg, mel, mask, dtype=torch.float64)
    margin = O.fsq_round_margin(z64, cfg.fsq_levels)
    model = _model(cfg, sd)
    idx = model.encode(mel.cuda(), mask.cuda())
    rep = index_report(idx, ref_idx, margin, TAU)
    REPORT["encode/hifispeech_4x200"] = rep
    print(rep)
    assert rep["safe_mismatch"] == 0 and rep["agree"] >= 0.995, rep
    ref = O.decode(w, cfg, ref_idx, mask, folded=True)
    out = model.decode(ref_idx.cuda(), mask.cuda()).cpu()
    err, scale = float((out - ref).abs().max()), float(ref.abs().max())
    REPORT["decode/hifispeech_4x200"] = {"max_abs_err": err, "ref_max": scale,
                                         "rel_l2": float((out - ref).norm() / ref.norm())}
    _dump()
    print(REPORT["decode/hifispeech_4x200"])
    assert err <= MEL_ATOL + MEL_RTOL * scale


@pytest.mark.parametrize("B,T,lengths", [
    (1, 1, [1]),                    # a single frame: every conv sees only padding around it
    (1, 5, [5]),                    # T < 8: the refiner pads to one 8-row block
    (2, 9, [9, 3]),                 # T % 8 == 1
    (3, 33, [33, 1, 17]),           # one-frame utterance inside a batch
    (2, 40, [40, 0]),               # an empty utterance (all padding) in the batch
    (1, 700, [700]),                # B = 1, many pair tiles along T
])
def test_edge_shapes_vs_oracle(B, T, lengths):
    """Ragged / tiny / empty inputs through the public API against the CPU oracle (tiny config)."""
    cfg, sd, _, _, _ = load_golden("tiny")
    lengths = torch.tensor(lengths)
    mel = synth_mels(B, T, cfg.mel_channels, seed=11)
    pad = torch.arange(T)[None, :] >= lengths[:, None]
    mel = mel.masked_fill(pad.unsqueeze(-1), 0.0)
    mask = pad.unsqueeze(1)
    w = O.effective_weights(sd)
    z32 = O.encode_latents(w, cfg, mel, mask, folded=True)
    ref_idx = O.fsq_quantize(z32, cfg.fsq_levels)[1]
    z64 = O.encode_latents(sd, cfg, mel, mask, dtype=torch.float64)
    margin = O.fsq_round_margin(z64, cfg.fsq_levels)
    model = _model(cfg, sd)
    idx = model.encode(mel.cuda(), mask.cuda())
    assert tuple(idx.shape) == (B, T)
    rep = index_report(idx, ref_idx, margin, TAU)
    assert rep["safe_mismatch"] == 0, rep
    ref = O.decode(w, cfg, ref_idx, mask, folded=True)
    out = model.decode(ref_idx.cuda(), mask.cuda()).cpu()
    assert out.shape == ref.shape
    err, scale = float((out - ref).abs().max()), float(ref.abs().max())
    assert err <= MEL_ATOL + MEL_RTOL * scale, (err, scale)
    assert bool(torch.isfinite(out).all())


def test_decode_rejects_out_of_range_indices():
    """bos/eos ids (codebook_size + 1/2, preencoder.py:340-341) are outside the FSQ range: the reference
    silently produces garbage digits; this path raises instead."""
    cfg, sd, mel, lengths, fx = load_golden("tiny")
    model = _model(cfg, sd)
    idx = torch.zeros(1, 16, dtype=torch.int64)
    idx[0, 3] = cfg.codebook_size + 1
    with pytest.raises((IndexError, RuntimeError)):
        model.decode(idx.cuda(), None)
    model.decode(torch.zeros(1, 16, dtype=torch.int64).cuda(), None)      # the engine is still usable afterwards


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_module_on_second_gpu_while_first_is_current():
    """The reference accepts get_pre_encoder(path, 'cuda:1') without a torch.cuda.set_device: every launch path here
    switches to the module's own device (ops.on_device), so results on cuda:1 equal those on cuda:0 bit for bit."""
    cfg, sd, mel, lengths, fx = load_golden("tiny")
    mask = sequence_mask(mel.shape[1], lengths).unsqueeze(1)
    torch.cuda.set_device(0)
    m0 = _model(cfg, sd)
    ref_idx = m0.encode(mel.cuda(0), mask.cuda(0))
    ref_out = m0.decode(ref_idx, mask.cuda(0))
    m1 = _model(cfg, sd).to("cuda:1")
    assert torch.cuda.current_device() == 0
    idx = m1.encode(mel.to("cuda:1"), mask.to("cuda:1"))
    out = m1.decode(idx, mask.to("cuda:1"))
    assert idx.device.index == 1 and out.device.index == 1 and torch.cuda.current_device() == 0
    assert torch.equal(idx.cpu(), ref_idx.cpu()) and torch.equal(out.cpu(), ref_out.cpu())
