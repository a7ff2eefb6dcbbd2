"""Per-kernel parity on the GPU, through the C ABI (ops.py -> libmqgan_b200.so).

Each kernel is compared with a float64 CPU restatement of the same reference op
on the same seeded inputs.  Tolerances are written beside each assert.
"""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from mqgan_b200 import ops  # noqa: E402
from oracle import preencoder_oracle as O  # noqa: E402

DEV = "cuda"


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def _ref_conv(x, w, b, kind):
    """x (N,H,W,C) float64 channel-last -> (N,H,W,Cout) float64."""
    N, H, W, Cc = x.shape
    if kind == "linear":
        return F.linear(x, w, b)
    if kind in ("same1d", "causal1d"):
        assert W == 1
        k = w.shape[2]
        xi = x.reshape(N, H, Cc).permute(0, 2, 1)
        if kind == "causal1d":
            y = F.conv1d(F.pad(xi, (k - 1, 0)), w, b)
        else:
            y = F.conv1d(xi, w, b, padding=(k - 1) // 2)
        return y.permute(0, 2, 1).reshape(N, H, 1, -1)
    if kind == "conv2d3":
        y = F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=1)
        return y.permute(0, 2, 3, 1)
    raise ValueError(kind)


CONV_CASES = [
    # kind, N, H, W, Cin, Cout, wshape-tail
    ("linear", 1, 256, 1, 128, 64, ()),
    ("linear", 2, 77, 1, 512, 144, ()),          # cout not a multiple of 32; ragged M
    ("same1d", 2, 300, 1, 64, 96, (3,)),
    ("same1d", 1, 130, 1, 192, 256, (5,)),
    ("causal1d", 2, 200, 1, 768, 512, (7,)),     # two N tiles, K = 7*768
    ("conv2d3", 2, 24, 144, 64, 128, (3, 3)),
    ("conv2d3", 1, 7, 36, 96, 192, (3, 3)),      # Cin not a multiple of 64 (hifimusic-like), small H
    ("conv2d3", 1, 16, 144, 192, 64, (3, 3)),
]


@pytest.mark.parametrize("kind,N,H,W,Cin,Cout,tail", CONV_CASES)
def test_conv_gemm_bf16(kind, N, H, W, Cin, Cout, tail):
    x = _rand(N, H, W, Cin, seed=1).to(torch.bfloat16)
    w = (_rand(Cout, Cin, *tail, seed=2) / (Cin * max(1, int(np.prod(tail)))) ** 0.5).to(torch.bfloat16)
    b = _rand(Cout, seed=3)
    ref = _ref_conv(x.double(), w.double(), b.double(), kind)
    pc = ops.pack_conv(w.float(), b, kind, split=False).to(DEV)
    out = torch.empty(N, H, W, Cout, dtype=torch.float32, device=DEV)
    ops.conv_gemm(x.to(DEV), pc, N, H, W, out_f32=out)
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item()
    # bf16 inputs are exact on both sides; only fp32 accumulation order differs
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), err


def test_conv_gemm_epilogue_variants():
    N, H, W, Cin, Cout = 2, 16, 48, 128, 128
    x = _rand(N, H, W, Cin, seed=4).to(torch.bfloat16)
    w = (_rand(Cout, Cin, 3, 3, seed=5) / (9 * Cin) ** 0.5).to(torch.bfloat16)
    b = _rand(Cout, seed=6)
    res = _rand(N, H, W, Cout, seed=7).to(torch.bfloat16)
    mask = torch.zeros(N, H, dtype=torch.uint8)
    mask[0, 11:] = 1
    mask[1, 5:] = 1
    acc = _ref_conv(x.double(), w.double(), b.double(), "conv2d3")
    m = mask.bool()[:, :, None, None]
    beta, gamma = 0.9, 0.6
    pc = ops.pack_conv(w.float(), b, "conv2d3", split=False).to(DEV)
    xd, rd, md = x.to(DEV), res.to(DEV), mask.to(DEV)

    # ConvBlock conv2: act -> + x -> mask (preencoder.py:98-101)
    ref = (O.aptx(acc, beta, gamma) + res.double()).masked_fill(m, 0.0)
    o32 = torch.empty(N, H, W, Cout, dtype=torch.float32, device=DEV)
    o16 = torch.empty(N, H, W, Cout, dtype=torch.bfloat16, device=DEV)
    ops.conv_gemm(xd, pc, N, H, W, row_mask=md, mask_post=True, act=True, beta=beta, gamma=gamma,
                  fast_tanh=False, res=rd, res_mode=2, out_f32=o32, out_bf16=o16)
    assert (o32.cpu().double() - ref).abs().max().item() < 1e-4
    assert (o16.cpu().double() - ref).abs().max().item() < 2e-2          # bf16 rounding of the output
    # fast tanh (tanh.approx): looser
    ops.conv_gemm(xd, pc, N, H, W, row_mask=md, mask_post=True, act=True, beta=beta, gamma=gamma,
                  fast_tanh=True, res=rd, res_mode=2, out_f32=o32)
    assert (o32.cpu().double() - ref).abs().max().item() < 5e-3

    # ResidualBlock1D tail: (+ res) -> mask -> act (attentions.py:545-549), fp32 residual, split output
    res32 = res.float()
    ref = O.aptx((acc + res32.double()).masked_fill(m, 0.0), beta, gamma)
    osp = torch.empty(N, H, W, 3 * Cout, dtype=torch.bfloat16, device=DEV)
    ops.conv_gemm(xd, pc, N, H, W, row_mask=md, mask_pre=True, act=True, beta=beta, gamma=gamma,
                  fast_tanh=False, res=res32.to(DEV), res_mode=1, out_f32=o32, out_split=osp)
    assert (o32.cpu().double() - ref).abs().max().item() < 1e-4
    s = osp.cpu().float().reshape(N, H, W, 3, Cout).sum(dim=3)
    assert (s - o32.cpu()).abs().max().item() < 1e-6                      # 3-term split carries fp32

    # strided fp32 output with a channel offset (out_proj / hidden_proj writing one buffer)
    big = torch.zeros(N, H, W, Cout + 32, dtype=torch.float32, device=DEV)
    ops.conv_gemm(xd, pc, N, H, W, out_f32=big, f32_coff=32)
    assert (big[..., 32:].cpu().double() - acc).abs().max().item() < 1e-4
    assert big[..., :32].abs().max().item() == 0.0


@pytest.mark.parametrize("kind,N,H,W,Cin,Cout,tail", [
    ("linear", 2, 100, 1, 128, 512, ()),
    ("same1d", 1, 257, 1, 512, 768, (5,)),
    ("linear", 1, 64, 1, 32, 64, ()),            # Cin < 64: segment over-read hits zero weights
])
def test_conv_gemm_bf16x3_is_fp32_grade(kind, N, H, W, Cin, Cout, tail):
    x = _rand(N, H, W, Cin, seed=8)
    w = _rand(Cout, Cin, *tail, seed=9) / (Cin * max(1, int(np.prod(tail)))) ** 0.5
    b = _rand(Cout, seed=10)
    ref = _ref_conv(x.double(), w.double(), b.double(), kind)
    pc = ops.pack_conv(w, b, kind, split=True).to(DEV)
    xs = ops.split_bf16(x.reshape(-1, Cin).to(DEV), 3)
    out = torch.empty(N, H, W, Cout, dtype=torch.float32, device=DEV)
    ops.conv_gemm(xs, pc, N, H, W, out_f32=out)
    err = (out.cpu().double() - ref).abs().max().item()
    ref32 = _ref_conv(x, w, b, kind)
    err32 = (ref32.double() - ref).abs().max().item()
    scale = max(1.0, ref.abs().max().item())
    # 24-bit operands; the remaining error is the tensor core's truncating fp32 accumulation
    # (~1e-5 relative at K = 2560), versus ~4e-3 for a single bf16 pass
    assert err < max(4 * err32, 2e-5 * scale), (err, err32)


@pytest.mark.parametrize("kind,N,H,W,Cin,Cout,tail", [
    ("linear", 2, 100, 1, 128, 512, ()),
    ("same1d", 1, 257, 1, 512, 768, (5,)),
    ("linear", 1, 64, 1, 32, 64, ()),
])
@pytest.mark.parametrize("wscale", [1.0, 1e-3, 300.0])
def test_conv_gemm_f16x2_is_fp32_grade(kind, N, H, W, Cin, Cout, tail, wscale):
    """Two fp16 terms per operand, three product segments: 22-bit operands at half the MMA work of
    bf16x3.  Weights are pre-scaled by a power of two at pack time (tiny weights would otherwise push
    the low term into fp16 subnormals) and the epilogue undoes it exactly."""
    x = _rand(N, H, W, Cin, seed=8) * 3.0
    w = _rand(Cout, Cin, *tail, seed=9) / (Cin * max(1, int(np.prod(tail)))) ** 0.5 * wscale
    b = _rand(Cout, seed=10) * wscale
    ref = _ref_conv(x.double(), w.double(), b.double(), kind)
    pc = ops.pack_conv(w, b, kind, split="f16x2").to(DEV)
    assert pc.wpack.dtype == torch.float16 and pc.nseg == 3
    xs = ops.split_bf16(x.reshape(-1, Cin).to(DEV), 2)
    assert xs.dtype == torch.float16
    h = xs.cpu().float().reshape(-1, 2, Cin)
    assert (h.sum(1) - x.reshape(-1, Cin)).abs().max().item() < 2e-6        # 22 bits of |x| <= 15
    out = torch.empty(N, H, W, Cout, dtype=torch.float32, device=DEV)
    osp = torch.empty(N, H, W, 2 * Cout, dtype=torch.float16, device=DEV)
    ops.conv_gemm(xs, pc, N, H, W, out_f32=out, out_split=osp)
    err = (out.cpu().double() - ref).abs().max().item()
    ref32 = _ref_conv(x, w, b, kind)
    err32 = (ref32.double() - ref).abs().max().item()
    scale = max(1e-30, ref.abs().max().item())
    assert err < max(4 * err32, 2e-5 * scale), (err, err32, scale)
    s2 = osp.cpu().float().reshape(N, H, W, 2, Cout).sum(dim=3)
    assert (s2 - out.cpu()).abs().max().item() < 1e-6 * max(1.0, scale) + 2e-7   # 2-term split of the output


def test_convblock2d_matches_reference_op():
    B, T, Cc = 2, 37, 96
    x = _rand(B, T, Cc, seed=11)
    w = {"p.dw.weight": _rand(1, 1, 5, 5, seed=12) * 0.2, "p.dw.bias": _rand(1, seed=13) * 0.1,
         "p.pw.weight": _rand(Cc, 1, 1, 1, seed=14), "p.pw.bias": _rand(Cc, seed=15),
         "p.conv_out.weight": _rand(1, Cc, 1, 1, seed=16) / Cc ** 0.5, "p.conv_out.bias": _rand(1, seed=17)}
    lengths = torch.tensor([T, 20])
    mask = O.sequence_mask(T, lengths)
    ref = O.convblock2d(x.permute(0, 2, 1).double(), mask.unsqueeze(1), {k: v.double() for k, v in w.items()}, "p")
    ref = ref.permute(0, 2, 1)
    dw = torch.cat([w["p.dw.weight"].reshape(25), w["p.dw.bias"]]).to(DEV)
    pw = torch.zeros(Cc, 4)
    pw[:, 0], pw[:, 1], pw[:, 2] = w["p.pw.weight"].reshape(Cc), w["p.pw.bias"], w["p.conv_out.weight"].reshape(Cc)
    m8 = mask.to(torch.uint8).to(DEV)
    o32 = torch.empty(B, T, Cc, dtype=torch.float32, device=DEV)
    osp = torch.empty(B, T, 3 * Cc, dtype=torch.bfloat16, device=DEV)
    ops.convblock2d(x.to(DEV), B, T, Cc, dw, pw.to(DEV), float(w["p.conv_out.bias"]), m8, False,
                    out_f32=o32, out_split=osp)
    err = (o32.cpu().double() - ref).abs().max().item()
    assert err < 5e-6 * max(1.0, ref.abs().max().item()), err
    assert (osp.cpu().float().reshape(B, T, 3, Cc).sum(2) - o32.cpu()).abs().max().item() < 1e-6
    # padded rows equal conv_out.bias exactly
    assert torch.all(o32[1, 20:] == float(w["p.conv_out.bias"]))
    # fast-tanh / bf16 input variant (decoder `post`)
    o16 = torch.empty(B, T, Cc, dtype=torch.bfloat16, device=DEV)
    xb = x.to(torch.bfloat16)
    refb = O.convblock2d(xb.permute(0, 2, 1).double(), mask.unsqueeze(1), {k: v.double() for k, v in w.items()}, "p").permute(0, 2, 1)
    ops.convblock2d(xb.to(DEV), B, T, Cc, dw, pw.to(DEV), float(w["p.conv_out.bias"]), m8, True, out_bf16=o16)
    assert (o16.cpu().double() - refb).abs().max().item() < 2e-2 * max(1.0, refb.abs().max().item())


def test_cbam_block_tail_matches_reference_op():
    B, T, Cc, R = 3, 150, 64, 8
    o = _rand(B, T, Cc, seed=20)
    r = _rand(B, T, Cc, seed=21)
    lengths = torch.tensor([150, 97, 31])
    mask = O.sequence_mask(T, lengths).unsqueeze(1)
    w = {"c.channel_attention.mlp.0.weight": _rand(R, Cc, seed=22) * 0.2, "c.channel_attention.mlp.0.bias": _rand(R, seed=23) * 0.1,
         "c.channel_attention.mlp.2.weight": _rand(Cc, R, seed=24) * 0.3, "c.channel_attention.mlp.2.bias": _rand(Cc, seed=25) * 0.1,
         "c.spatial_attention.conv.weight": _rand(1, 2, 7, seed=26) * 0.3}
    wd = {k: v.double() for k, v in w.items()}
    beta, gamma = 1.1, 0.45
    cb = O.cbam(o.permute(0, 2, 1).double(), mask, wd, "c")
    ref = O.aptx((cb + r.permute(0, 2, 1).double()).masked_fill(mask, 0), beta, gamma).permute(0, 2, 1)
    m8 = mask.squeeze(1).to(torch.uint8).to(DEV)
    od = o.to(DEV)
    gate = ops.cam_gate(od, m8, B, T, Cc, *(w[k].to(DEV).contiguous() for k in list(w)[:4]))
    y = torch.empty(B, T, Cc, dtype=torch.float32, device=DEV)
    ops.cbam_apply(od, gate, r.to(DEV), m8, B, T, Cc, w["c.spatial_attention.conv.weight"].reshape(14).to(DEV),
                   beta, gamma, out_f32=y)
    err = (y.cpu().double() - ref).abs().max().item()
    assert err < 5e-6 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("levels", [[8, 5, 5, 5], [8, 8, 5, 5, 5]])
def test_fsq_bit_exact_vs_reference_vectors(levels, golden_dir):
    fx = np.load(os.path.join(golden_dir, "fsq_" + "_".join(map(str, levels)) + ".npz"))
    z = torch.from_numpy(fx["z"]).reshape(-1, len(levels)).contiguous()
    fsq = ops.fsq_params(levels)
    idx, codes = ops.fsq_quantize(z.to(DEV), fsq, want_codes=True)
    ref_idx = torch.from_numpy(fx["indices"].astype(np.int64)).reshape(-1)
    # integer work: bit-exact except where CUDA tanhf and the host tanh differ by an ulp exactly at a
    # rounding boundary; the fixture is random so demand exact equality
    assert torch.equal(idx.cpu(), ref_idx)
    assert torch.equal(codes.cpu(), torch.from_numpy(fx["codes"]).reshape(-1, len(levels)))
    # adversarial: exact half-way points in bounded space round half-to-even like torch.round
    lv, basis, half_l, offset, shift, half_w = O.fsq_constants(levels)
    n = int(np.prod(levels))
    table = torch.from_numpy(fx["all_codes"])
    zero = torch.zeros(1, len(levels))
    i0 = ops.fsq_quantize(zero.to(DEV), fsq)
    assert int(i0) == int(O.fsq_quantize(zero, levels)[1])


@pytest.mark.parametrize("levels", [[8, 5, 5, 5], [8, 8, 5, 5, 5]])
def test_fsq_adversarial_and_million_vs_reference(levels, golden_dir):
    """The reference's own FSQ.forward answers (quantizer.py:109-140,177-181) on (a) latents whose bounded image sits on
    a rounding boundary k + 0.5 and their +-1..3 ulp neighbours, saturating magnitudes, signed zeros, (b) 2^20 random
    latents.  Integer work: equal everywhere except where the float64 bounded value is within 1e-6 of a boundary -
    there the answer depends on the last ulp of tanh (CUDA tanhf vs the host's vectorised tanh), which no
    implementation can promise; those rows are counted and must stay a handful."""
    fx = np.load(os.path.join(golden_dir, "fsq_" + "_".join(map(str, levels)) + ".npz"))
    fsq = ops.fsq_params(levels)
    rep = {}
    g = torch.Generator().manual_seed(int(fx["big_seed"]))
    z_big = (torch.randn(1, int(fx["big_n"]), len(levels), generator=g) * float(fx["big_scale"]))[0].contiguous()
    for tag, z, ref in (("adversarial", torch.from_numpy(fx["z_adv"]).contiguous(), fx["indices_adv"]),
                        ("million", z_big, fx["indices_big"])):
        ref = torch.from_numpy(ref.astype(np.int64))
        idx = ops.fsq_quantize(z.to(DEV), fsq).cpu()
        margin = O.fsq_round_margin(z.double(), levels)
        neq = idx != ref
        rep[tag] = {"rows": int(z.shape[0]), "mismatch": int(neq.sum()), "mismatch_outside_1e-6": int((neq & (margin > 1e-6)).sum()),
                    "rows_within_1e-6_of_a_boundary": int((margin <= 1e-6).sum())}
        assert rep[tag]["mismatch_outside_1e-6"] == 0, rep
        assert int(idx.min()) >= 0 and int(idx.max()) < int(np.prod(levels))
    print("fsq", levels, rep)
    assert rep["million"]["mismatch"] <= 8, rep                      # measured: 0-2 of 2^20
    # the fused projection + quantiser kernel sees the same latents through an identity projection
    D = len(levels)
    eye = torch.zeros(D, 8)
    eye[:, :D] = torch.eye(D)
    y = torch.zeros(z_big.shape[0], 8)
    y[:, :D] = z_big
    idx2 = ops.qin_fsq(y.to(DEV).contiguous(), eye.to(DEV).contiguous(), torch.zeros(D, device=DEV), fsq).cpu()
    ref = torch.from_numpy(fx["indices_big"].astype(np.int64))
    margin = O.fsq_round_margin(z_big.double(), levels)
    assert int(((idx2 != ref) & (margin > 1e-6)).sum()) == 0


def test_qin_fsq_and_gather():
    levels = [8, 5, 5, 5]
    rows, Cc = 1000, 768
    y = _rand(rows, Cc, seed=30)
    w = _rand(4, Cc, seed=31) / Cc ** 0.5
    b = _rand(4, seed=32) * 0.1
    fsq = ops.fsq_params(levels)
    idx, z = ops.qin_fsq(y.to(DEV), w.to(DEV), b.to(DEV), fsq, want_z=True)
    z64 = F.linear(y.double(), w.double(), b.double())
    assert (z.cpu().double() - z64).abs().max().item() < 1e-6
    ref_idx = O.fsq_quantize(z.cpu(), levels)[1]
    assert torch.equal(idx.cpu(), ref_idx)
    # K8: gather == indices_to_codes + q_out_proj
    wq, bq = _rand(Cc, 4, seed=33), _rand(Cc, seed=34)
    codes = O.fsq_indices_to_codes(torch.arange(1000), levels)
    table = F.linear(codes, wq, bq).contiguous()
    ob, of = ops.code_gather(idx, table.to(DEV), bf16=True, f32=True)
    ref = F.linear(O.fsq_indices_to_codes(idx.cpu(), levels), wq, bq)
    assert torch.equal(of.cpu(), ref)
    assert torch.equal(ob.cpu(), ref.to(torch.bfloat16))
    bad = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.code_gather(torch.tensor([3, 1001, -1], device=DEV), table.to(DEV), bad=bad)
    assert int(bad) == 1                                 # bos/eos ids are outside the FSQ range (App. B7)


def test_refiner_masks_pool_upcat_stem_tail():
    B, T, depth, Fw, Cc, M = 2, 21, 3, 20, 16, 16
    lengths = torch.tensor([21, 13])
    mask = O.sequence_mask(T, lengths)
    T8, down, up = ops.refiner_masks(mask.to(torch.uint8).to(DEV), B, T, depth, DEV)
    assert T8 == 24
    cur = torch.cat([mask, torch.ones(B, T8 - T, dtype=torch.bool)], 1).reshape(B, 1, T8, 1)
    downs_ref = [cur]
    for _ in range(depth):
        cur = F.max_pool2d(cur.float(), kernel_size=(2, 1), stride=(2, 1)).bool()
        downs_ref.append(cur)
    ups_ref = {depth: cur}
    for l in range(depth - 1, -1, -1):
        cur = F.interpolate(cur.float(), scale_factor=(2, 1), mode="nearest").bool()
        ups_ref[l] = cur
    for l in range(depth + 1):
        assert torch.equal(down[l].cpu().bool(), downs_ref[l].reshape(B, -1))
        assert torch.equal(up[l].cpu().bool(), ups_ref[l].reshape(B, -1))
    # no mask -> only the T padding is masked
    _, d0, _ = ops.refiner_masks(None, B, T, depth, DEV)
    assert d0[0][:, :T].sum().item() == 0 and d0[0][:, T:].all()

    x = _rand(B, T8, Fw, Cc, seed=40).to(torch.bfloat16)
    y = ops.avgpool_mask(x.to(DEV), down[1], B, T8, Fw, Cc)
    ref = F.avg_pool2d(x.float().permute(0, 3, 1, 2), kernel_size=(2, 1)).masked_fill(downs_ref[1], 0.0).permute(0, 2, 3, 1)
    assert torch.equal(y.cpu(), ref.to(torch.bfloat16))
    lo = _rand(B, T8 // 2, Fw, 2 * Cc, seed=41).to(torch.bfloat16)
    u = ops.upcat_mask(lo.to(DEV), x.to(DEV), up[0], B, T8, Fw, 2 * Cc, Cc)
    ref = torch.cat([F.interpolate(lo.float().permute(0, 3, 1, 2), scale_factor=(2, 1), mode="nearest"),
                     x.float().permute(0, 3, 1, 2)], 1).masked_fill(ups_ref[0], 0.0).permute(0, 2, 3, 1)
    assert torch.equal(u.cpu(), ref.to(torch.bfloat16))

    # stem: conv 1->C over the masked, T-padded image + APTx
    r = _rand(B, T, Fw, seed=42)
    w1, b1 = _rand(Cc, 1, 3, 3, seed=43) * 0.3, _rand(Cc, seed=44) * 0.1
    img = torch.cat([r, torch.zeros(B, T8 - T, Fw)], 1).unsqueeze(1).masked_fill(downs_ref[0], 0.0)
    ref = O.aptx(F.conv2d(img.double(), w1.double(), b1.double(), padding=1), 1, 0.5).permute(0, 2, 3, 1)
    s = ops.refiner_stem(r.to(DEV), mask.to(torch.uint8).to(DEV), B, T, T8, Fw, Cc,
                         w1.reshape(Cc, 9).contiguous().to(DEV), b1.to(DEV), False)
    assert (s.cpu().double() - ref).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())

    # tail: conv C->1 + crop + mask + reproj + x_recon add
    xin = _rand(B, T8, Fw, Cc, seed=45).to(torch.bfloat16)
    wp, bp = _rand(1, Cc, 3, 3, seed=46) / (9 * Cc) ** 0.5, _rand(1, seed=47)
    wr = _rand(M, Fw, seed=48) / Fw ** 0.5
    o = F.conv2d(xin.double().permute(0, 3, 1, 2), wp.double(), bp.double(), padding=1).squeeze(1)[:, :T]
    o = o.masked_fill(mask.unsqueeze(-1), 0.0)
    ref = r[..., :M].double() + F.linear(o, wr.double())
    pc9 = ops.pack_conv(wp.reshape(Cc, 9).t().contiguous(), None, "linear", False).to(DEV)
    tp = torch.empty(B, T8, Fw, 12, dtype=torch.float32, device=DEV)
    ops.conv_gemm(xin.to(DEV), pc9, B, T8, Fw, out_f32=tp)
    out = ops.refiner_tail(tp, mask.to(torch.uint8).to(DEV), B, T, T8, Fw, float(bp), wr.t().contiguous().to(DEV), M,
                           r.to(DEV))
    # the 9 tap weights are bf16 inside the GEMM (like every decoder weight): bf16-level tolerance
    assert (out.cpu().double() - ref).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())


def test_sequence_mask_kernel():
    lengths = torch.tensor([5, 0, 9, 3])
    m = ops.sequence_mask(lengths.to(DEV), 9)
    assert torch.equal(m.cpu().bool(), O.sequence_mask(9, lengths))


def test_convblock2d_table_mode_matches_exact_sum():
    """The tabulated g(s) (per-interval cubics, built in float64) against the float64 oracle and
    the exact-sum kernel, including pixels outside the table range (exact fallback)."""
    from mqgan_b200.engine import build_cb2d_table
    B, T, Cc = 2, 45, 128
    x = _rand(B, T, Cc, seed=50) * 3.0
    x[0, 7, 40:48] = 400.0           # pushes some s beyond +-64 -> exact fallback path
    w = {"p.dw.weight": _rand(1, 1, 5, 5, seed=51) * 0.3, "p.dw.bias": _rand(1, seed=52) * 0.1,
         "p.pw.weight": _rand(Cc, 1, 1, 1, seed=53), "p.pw.bias": _rand(Cc, seed=54),
         "p.conv_out.weight": _rand(1, Cc, 1, 1, seed=55) / Cc ** 0.5, "p.conv_out.bias": _rand(1, seed=56)}
    lengths = torch.tensor([T, 30])
    mask = O.sequence_mask(T, lengths)
    ref = O.convblock2d(x.permute(0, 2, 1).double(), mask.unsqueeze(1), {k: v.double() for k, v in w.items()}, "p").permute(0, 2, 1)
    dw = torch.cat([w["p.dw.weight"].reshape(25), w["p.dw.bias"]]).to(DEV)
    pw = torch.zeros(Cc, 4)
    pw[:, 0], pw[:, 1], pw[:, 2] = w["p.pw.weight"].reshape(Cc), w["p.pw.bias"], w["p.conv_out.weight"].reshape(Cc)
    bout = float(w["p.conv_out.bias"])
    coef, off, inv_h, terr = build_cb2d_table(pw[:, 0], pw[:, 1], pw[:, 2], bout, device=DEV)
    assert terr <= 2.5e-7
    m8 = mask.to(torch.uint8).to(DEV)
    ot = torch.empty(B, T, Cc, dtype=torch.float32, device=DEV)
    oe = torch.empty(B, T, Cc, dtype=torch.float32, device=DEV)
    osp = torch.empty(B, T, 3 * Cc, dtype=torch.bfloat16, device=DEV)
    ops.convblock2d(x.to(DEV), B, T, Cc, dw, pw.to(DEV), bout, m8, False, out_f32=ot, out_split=osp,
                    table=coef.to(DEV), table_off=off, table_inv_h=inv_h)
    ops.convblock2d(x.to(DEV), B, T, Cc, dw, pw.to(DEV), bout, m8, False, out_f32=oe)
    scale = max(1.0, ref.abs().max().item())
    e_t = (ot.cpu().double() - ref).abs().max().item()
    e_e = (oe.cpu().double() - ref).abs().max().item()
    print("table err", e_t, "exact-kernel err", e_e, "scale", scale)
    assert e_t < 5e-6 * scale and e_e < 5e-6 * scale
    assert (osp.cpu().float().reshape(B, T, 3, Cc).sum(2) - ot.cpu()).abs().max().item() < 1e-6 * scale
    assert torch.all(ot[1, 30:] == bout)


@pytest.mark.parametrize("kind,N,H,W,Cin,Cout,tail,msub", [
    ("conv2d3", 2, 40, 144, 64, 64, (3, 3), 4),      # narrow layer: 4 sub-tiles share one weight tile
    ("conv2d3", 1, 21, 36, 128, 128, (3, 3), 2),     # H not a multiple of msub*bh
    ("same1d", 2, 700, 1, 64, 96, (3,), 2),
    ("causal1d", 1, 513, 1, 128, 32, (5,), 4),
])
def test_conv_gemm_multi_subtile(kind, N, H, W, Cin, Cout, tail, msub):
    x = _rand(N, H, W, Cin, seed=61).to(torch.bfloat16)
    w = (_rand(Cout, Cin, *tail, seed=62) / (Cin * max(1, int(np.prod(tail)))) ** 0.5).to(torch.bfloat16)
    b = _rand(Cout, seed=63)
    mask = torch.zeros(N, H, dtype=torch.uint8)
    mask[0, H // 2:] = 1
    ref = O.aptx(_ref_conv(x.double(), w.double(), b.double(), kind), 1.0, 0.5)
    ref = ref.masked_fill(mask.bool()[:, :, None, None], 0.0)
    pc = ops.pack_conv(w.float(), b, kind, split=False).to(DEV)
    out = torch.empty(N, H, W, Cout, dtype=torch.float32, device=DEV)
    ops.conv_gemm(x.to(DEV), pc, N, H, W, row_mask=mask.to(DEV), mask_post=True, act=True, fast_tanh=False,
                  out_f32=out, msub=msub, pair=False)
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("N,Hl,W,Cx,Cs,Cout,msub,pair", [
    (2, 8, 48, 128, 64, 64, None, False),
    (1, 11, 36, 96, 96, 96, 1, False),          # odd half-rows, channel counts that are not multiples of 64
    (2, 20, 144, 256, 128, 128, 2, False),
    # CTA-pair main loop (cta_group::2): x halo + two skip-parity boxes per chunk
    (2, 32, 48, 128, 64, 64, 1, True),
    (1, 35, 36, 96, 96, 96, 1, True),           # ragged rows / columns, bn = 96 (48 weight rows per CTA)
    (2, 72, 144, 256, 128, 128, 2, True),       # two sub-tiles per CTA, last pair tile half empty
    (1, 64, 144, 512, 256, 256, None, True),    # ups.0.conv1 shape, automatic msub
    (2, 8, 24, 64, 64, 64, 1, True),            # image smaller than one pair tile
])
def test_conv_gemm_fused_upsample_concat(N, Hl, W, Cx, Cs, Cout, msub, pair):
    """UpBlock: conv3x3(mask(cat[nearest_up(x), skip])) without materialising the concat."""
    H = 2 * Hl
    x = _rand(N, Hl, W, Cx, seed=70).to(torch.bfloat16)
    skip = _rand(N, H, W, Cs, seed=71).to(torch.bfloat16)
    w = (_rand(Cout, Cx + Cs, 3, 3, seed=72) / (9 * (Cx + Cs)) ** 0.5).to(torch.bfloat16)
    b = _rand(Cout, seed=73)
    mask_new = torch.zeros(N, H, dtype=torch.uint8)
    mask_new[0, H - 4:] = 1
    mask_old = torch.zeros(N, H, dtype=torch.uint8)
    mask_old[0, H - 2:] = 1
    x = x.masked_fill(mask_new[:, ::2].bool()[:, :, None, None], 0)      # low-res input is already masked
    skip = skip.masked_fill(mask_old.bool()[:, :, None, None], 0)        # skip carries the finer down-path mask
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=(2, 1), mode="nearest")
    cat = torch.cat([up, skip.float().permute(0, 3, 1, 2)], 1).masked_fill(mask_new.bool()[:, None, :, None], 0.0)
    ref = O.aptx(F.conv2d(cat.double(), w.double(), b.double(), padding=1), 1.0, 0.5).permute(0, 2, 3, 1)
    pc = ops.pack_upconv(w.float(), b, Cx, Cs).to(DEV)
    sd = skip.to(DEV)
    ops.zero_rows(sd, mask_new.to(DEV), mask_old.to(DEV))
    out = torch.empty(N, H, W, Cout, dtype=torch.bfloat16, device=DEV)
    ops.conv_gemm(x.to(DEV), pc, N, Hl, W, x2=sd, act=True, fast_tanh=False, out_bf16=out, msub=msub, pair=pair)
    err = (out.cpu().double() - ref).abs().max().item()
    # pre-summed tap weights are rounded to bf16 once more; output is bf16
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), err
    # masked-output variant with fp32 output for a tighter check of the indexing
    o32 = torch.empty(N, H, W, Cout, dtype=torch.float32, device=DEV)
    ops.conv_gemm(x.to(DEV), pc, N, Hl, W, x2=sd, act=True, fast_tanh=False, row_mask=mask_new.to(DEV),
                  mask_post=True, out_f32=o32, msub=msub, pair=pair)
    ref_m = ref.masked_fill(mask_new.bool()[:, :, None, None], 0.0)
    assert (o32.cpu().double() - ref_m).abs().max().item() < 6e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("N,H,W,Cin,Cout,msub", [
    (2, 40, 144, 64, 64, 4),
    (1, 37, 36, 128, 128, 2),        # H not a multiple of the tile; W = 36 -> 5 column tiles (last one ragged)
    (2, 16, 144, 256, 256, 1),
    (1, 9, 20, 96, 192, 1),          # Cin padded to 128, tiny image
    (3, 128, 144, 192, 64, None),    # automatic msub
    (2, 64, 144, 128, 256, 2),       # bn = 256 with two sub-tiles: single TMEM accumulator buffer
    (1, 40, 72, 64, 512, 2),         # two N tiles, single-buffered
])
def test_conv_halo_mode_matches_tap_mode(N, H, W, Cin, Cout, msub):
    """Halo-tile main loop (one activation fetch, 9 shifted descriptors) against the float64 conv and
    against the tap-shifted main loop (must agree bit for bit: same K order per output)."""
    x = _rand(N, H, W, Cin, seed=81).to(torch.bfloat16)
    w = (_rand(Cout, Cin, 3, 3, seed=82) / (9 * Cin) ** 0.5).to(torch.bfloat16)
    b = _rand(Cout, seed=83)
    res = _rand(N, H, W, Cout, seed=84).to(torch.bfloat16)
    mask = torch.zeros(N, H, dtype=torch.uint8)
    mask[0, H // 3:] = 1
    ref = (O.aptx(_ref_conv(x.double(), w.double(), b.double(), "conv2d3"), 1.0, 0.5) + res.double())
    ref = ref.masked_fill(mask.bool()[:, :, None, None], 0.0)
    pc = ops.pack_conv(w.float(), b, "conv2d3", split=False).to(DEV)
    outs = {}
    for halo in (True, False):
        o = torch.empty(N, H, W, Cout, dtype=torch.float32, device=DEV)
        ops.conv_gemm(x.to(DEV), pc, N, H, W, row_mask=mask.to(DEV), mask_post=True, act=True, fast_tanh=False,
                      res=res.to(DEV), res_mode=2, out_f32=o, msub=msub, halo=halo, pair=False)
        outs[halo] = o.cpu()
    err = (outs[True].double() - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), err
    # tap-mode accumulates (tap, chunk) and halo-mode (chunk, tap): tiny fp32 reordering differences only
    assert (outs[True] - outs[False]).abs().max().item() < 1e-4
    ob = torch.empty(N, H, W, Cout, dtype=torch.bfloat16, device=DEV)
    ops.conv_gemm(x.to(DEV), pc, N, H, W, row_mask=mask.to(DEV), mask_post=True, act=True, out_bf16=ob, halo=True,
                  msub=msub, pair=False)
    ref2 = O.aptx(_ref_conv(x.double(), w.double(), b.double(), "conv2d3"), 1.0, 0.5).masked_fill(mask.bool()[:, :, None, None], 0.0)
    assert (ob.cpu().double() - ref2).abs().max().item() < 2e-2 * max(1.0, ref2.abs().max().item())


@pytest.mark.parametrize("N,H,W,Cin,Cout,msub", [
    (2, 128, 144, 64, 64, 4),        # narrow layer: 4 sub-tiles per CTA, 8 per pair
    (1, 70, 36, 128, 128, 2),        # H not a multiple of the pair tile; W = 36 -> ragged last column tile
    (2, 32, 144, 256, 256, 1),       # exactly one pair tile of rows
    (1, 40, 72, 64, 512, 2),         # two N tiles, single TMEM accumulator buffer (msub*bn = 512)
    (1, 33, 20, 96, 192, 1),         # Cin padded to 128, bn = 192 (96 weight rows per CTA)
    (1, 64, 40, 96, 96, 2),          # hifimusic width: bn = 96 (48 weight rows per CTA)
    (3, 128, 144, 192, 64, None),    # automatic msub
    (1, 20, 16, 64, 64, 1),          # image smaller than one pair tile: the peer CTA works on padding only
    (4, 128, 144, 512, 512, None),   # mid.conv shape: many pair tiles per cluster (ring wrap-around)
])
def test_conv_pair_mode_matches_tap_mode(N, H, W, Cin, Cout, msub):
    """CTA-pair main loop (tcgen05 cta_group::2, M = 256, half a weight tile per CTA) against the
    float64 convolution and against the single-CTA tap-shifted main loop."""
    x = _rand(N, H, W, Cin, seed=91).to(torch.bfloat16)
    w = (_rand(Cout, Cin, 3, 3, seed=92) / (9 * Cin) ** 0.5).to(torch.bfloat16)
    b = _rand(Cout, seed=93)
    res = _rand(N, H, W, Cout, seed=94).to(torch.bfloat16)
    mask = torch.zeros(N, H, dtype=torch.uint8)
    mask[0, H // 3:] = 1
    ref = (O.aptx(_ref_conv(x.double(), w.double(), b.double(), "conv2d3"), 1.0, 0.5) + res.double())
    ref = ref.masked_fill(mask.bool()[:, :, None, None], 0.0)
    pc = ops.pack_conv(w.float(), b, "conv2d3", split=False).to(DEV)
    outs = {}
    for pair in (True, False):
        o = torch.empty(N, H, W, Cout, dtype=torch.float32, device=DEV)
        ops.conv_gemm(x.to(DEV), pc, N, H, W, row_mask=mask.to(DEV), mask_post=True, act=True, fast_tanh=False,
                      res=res.to(DEV), res_mode=2, out_f32=o, msub=msub if pair else None, pair=pair, halo=False)
        outs[pair] = o.cpu()
    err = (outs[True].double() - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), err
    assert (outs[True] - outs[False]).abs().max().item() < 1e-4      # fp32 reordering differences only
    # lean bf16 epilogue, run twice: persistent ring state must not leak between launches
    ref2 = O.aptx(_ref_conv(x.double(), w.double(), b.double(), "conv2d3"), 1.0, 0.5).masked_fill(mask.bool()[:, :, None, None], 0.0)
    for _ in range(2):
        ob = torch.zeros(N, H, W, Cout, dtype=torch.bfloat16, device=DEV)
        ops.conv_gemm(x.to(DEV), pc, N, H, W, row_mask=mask.to(DEV), mask_post=True, act=True, out_bf16=ob, pair=True,
                      msub=msub)
        assert (ob.cpu().double() - ref2).abs().max().item() < 2e-2 * max(1.0, ref2.abs().max().item())


# ---------------------------------------------------------------------------
# nearest-codeword lookup (BASELINE configs[3]): distance GEMM + argmin + gather
# ---------------------------------------------------------------------------
def _vq_ref(z, cb):
    """float64 argmin of ||z - c||^2 (lowest index on ties) and the gap to the runner-up."""
    d = torch.cdist(z.double(), cb.double()) ** 2
    best2 = torch.topk(d, 2, dim=1, largest=False)
    idx = torch.argmin(d, dim=1)          # first occurrence = lowest index
    return idx, best2.values[:, 1] - best2.values[:, 0], d


@pytest.mark.parametrize("levels", [[8, 8, 4, 4], [8, 8, 8, 4, 4], [8, 5, 5, 5]])
def test_vq_nearest_equals_reference_fsq_on_implicit_codebook(levels):
    """On FSQ's implicit codebook (quantizer.py:101-104) the generic nearest-codeword kernel must give
    the reference quantiser's index (quantizer.py:128-140,177-181) wherever the latent is not within
    1e-4 of a rounding boundary, and gather exactly the reference's codes."""
    D, K = len(levels), int(np.prod(levels))
    codebook = O.fsq_indices_to_codes(torch.arange(K), levels)                     # (K, D), the implicit codebook
    z = _rand(20000, D, seed=40) * 1.5
    codes_ref, idx_ref = O.fsq_quantize(z, levels)
    lv, basis, half_l, offset, shift, half_w = O.fsq_constants(levels)
    y = (((z + shift).tanh() * half_l - offset) / half_w).contiguous()             # bounded latent in code units
    margin = O.fsq_round_margin(z, levels)
    pc = ops.pack_codebook(codebook, "f16x2").to(DEV)
    assert pc.k == K and pc.k_pad % 256 == 0
    idx, codes = ops.vq_nearest(y.to(DEV), pc, want_codes=True)
    neq = idx.cpu() != idx_ref
    assert int((neq & (margin > 1e-4)).sum()) == 0, int(neq.sum())
    assert float(neq.float().mean()) < 1e-3
    assert torch.equal(codes.cpu(), codebook[idx.cpu()])                          # gather is exact
    assert torch.equal(codes.cpu()[~neq], codes_ref[~neq])


@pytest.mark.parametrize("n,K,D", [
    (5000, 1024, 64),      # one row tile is ragged (5000 = 39*128 + 8)
    (3000, 1000, 64),      # K not a multiple of the 256-code accumulator block: padding codes carry +inf
    (2048, 8192, 64),      # 32 accumulator blocks per row tile: ||c||^2 exactly fills its shared-memory stage
    (1500, 700, 20),       # two K-steps per code, two codes per 128-byte row
    (1500, 3000, 40),      # D padded to 64
    (4097, 512, 4),        # four codes per row
    (100, 300, 1),
])
def test_vq_nearest_generic_codebook(n, K, D):
    z = _rand(n, D, seed=41)
    cb = _rand(K, D, seed=42)
    ref_idx, gap, d64 = _vq_ref(z, cb)
    scale = float(d64.min(dim=1).values.mean()) + 1.0
    pc = ops.pack_codebook(cb, "f16x2").to(DEV)
    idx, codes, dist = ops.vq_nearest(z.to(DEV), pc, want_codes=True, want_dist=True)
    idx = idx.cpu()
    neq = idx != ref_idx
    # fp32-grade: disagreement only where the float64 runner-up is within 1e-5 (relative) of the winner
    assert int((neq & (gap > 1e-5 * scale)).sum()) == 0, (int(neq.sum()), float(gap[neq].max()) if neq.any() else 0)
    assert int(idx.min()) >= 0 and int(idx.max()) < K
    assert torch.equal(codes.cpu(), cb[idx])
    # dist_out = ||c||^2 - 2 z.c at the minimum
    ref_d = (d64.gather(1, idx[:, None]).squeeze(1) - (z.double() ** 2).sum(1))
    assert (dist.cpu().double() - ref_d).abs().max().item() < 1e-4 * scale
    # bf16 mode: agreement rate is reported, not exact
    pcb = ops.pack_codebook(cb, "bf16").to(DEV)
    ib = ops.vq_nearest(z.to(DEV), pcb, want_codes=False).cpu()
    agree = float((ib == ref_idx).float().mean())
    print(f"vq bf16 agreement n={n} K={K} D={D}: {agree:.4f}")
    if D >= 20:
        assert agree > 0.80


def test_vq_nearest_ties_pick_lowest_index():
    D = 8
    base = _rand(300, D, seed=43)
    cb = torch.cat([base, base, base[:100]])          # every code appears 2-3 times
    z = base[torch.randint(0, 300, (1000,), generator=torch.Generator().manual_seed(44))] + 0.01 * _rand(1000, D, seed=45)
    pc = ops.pack_codebook(cb, "f16x2").to(DEV)
    idx = ops.vq_nearest(z.to(DEV), pc, want_codes=False).cpu()
    ref_idx, _, _ = _vq_ref(z, cb)
    assert int(idx.max()) < 300                         # duplicates at k + 300 / k + 600 never win
    assert float((idx == ref_idx).float().mean()) > 0.999


@pytest.mark.parametrize("N,H,W,Cin,Cout,pair", [
    (2, 64, 144, 64, 64, True),          # pre.conv2-like: 4 sub-tiles per CTA
    (1, 70, 36, 128, 128, True),         # ragged rows (70 = 2*32 + 6) and columns
    (2, 32, 144, 256, 256, True),
    (1, 22, 20, 64, 128, False),         # single-CTA halo loop
])
def test_conv_epilogue_fused_avgpool_equals_separate_pass(N, H, W, Cin, Cout, pair):
    """DownBlock's AvgPool2d((2,1)) + pooled-mask fill written by the producing conv's epilogue must
    equal mq_avgpool_mask applied to the stored bf16 output bit for bit (preencoder.py:111-114)."""
    x = _rand(N, H, W, Cin, seed=101).to(torch.bfloat16)
    w = (_rand(Cout, Cin, 3, 3, seed=102) / (9 * Cin) ** 0.5).to(torch.bfloat16)
    b = _rand(Cout, seed=103)
    res = _rand(N, H, W, Cout, seed=104).to(torch.bfloat16) if Cin == Cout else None
    mask = torch.zeros(N, H, dtype=torch.uint8)
    mask[0, H // 2 + 1:] = 1                                   # odd boundary: one pooling pair is half padded
    mask_pool = (mask.view(N, H // 2, 2).max(dim=2).values).contiguous()
    pc = ops.pack_conv(w.float(), b, "conv2d3", split=False).to(DEV)
    y = torch.empty(N, H, W, Cout, dtype=torch.bfloat16, device=DEV)
    yp = torch.full((N, H // 2, W, Cout), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.conv_gemm(x.to(DEV), pc, N, H, W, row_mask=mask.to(DEV), mask_post=True, act=True,
                  res=None if res is None else res.to(DEV), res_mode=0 if res is None else 2,
                  out_bf16=y, out_pool=yp, pair=pair, halo=not pair)
    y2 = torch.empty_like(y)
    ops.conv_gemm(x.to(DEV), pc, N, H, W, row_mask=mask.to(DEV), mask_post=True, act=True,
                  res=None if res is None else res.to(DEV), res_mode=0 if res is None else 2,
                  out_bf16=y2, pair=pair, halo=not pair)
    assert torch.equal(y, y2)                                   # the pooled output does not disturb the main one
    ref = ops.avgpool_mask(y2, mask_pool.to(DEV), N, H, W, Cout)
    assert torch.equal(yp, ref)
    # and against the float64 definition
    r64 = y2.cpu().double().view(N, H // 2, 2, W, Cout).mean(dim=2).masked_fill(mask_pool.bool()[:, :, None, None], 0.0)
    assert (yp.cpu().double() - r64).abs().max().item() < 1e-2 * max(1.0, r64.abs().max().item())


@pytest.mark.parametrize("kind,N,H,Cin,Cout,tail,mode", [
    ("same1d", 2, 1000, 512, 512, (3,), "bf16"),       # 3 taps = one weight-ring slot; H not a multiple of 256
    ("same1d", 1, 700, 512, 768, (5,), "f16x2"),       # 5 taps = ragged slot groups (3 + 2), three N tiles, 3 segments
    ("causal1d", 2, 512, 768, 512, (7,), "bf16"),      # 7 taps (3 + 3 + 1), left halo only
    ("causal1d", 1, 300, 512, 384, (5,), "bf16"),      # bn = 192 (96 weight rows per CTA), one ragged pair tile
    ("linear", 3, 256, 128, 512, (), "f16x2"),         # a single tap
    ("same1d", 1, 1024, 384, 512, (3,), "bf16x3"),     # six segments
])
def test_conv_pair_1d_matches_tap_mode(kind, N, H, Cin, Cout, tail, mode):
    """Row-halo CTA-pair loop for 1-D convolutions against float64 and the single-CTA tap-shifted loop."""
    x = _rand(N, H, 1, Cin, seed=111)
    w = _rand(Cout, Cin, *tail, seed=112) / (Cin * max(1, int(np.prod(tail)))) ** 0.5
    b = _rand(Cout, seed=113)
    res = _rand(N, H, 1, Cout, seed=114)
    mask = torch.zeros(N, H, dtype=torch.uint8)
    mask[0, H // 2:] = 1
    m = mask.bool()[:, :, None, None]
    if mode == "bf16":
        xq, wq = x.to(torch.bfloat16), w.to(torch.bfloat16)
        xin = xq.to(DEV)
        ref_in, ref_w = xq.double(), wq.double()
        tol = 2e-4
    else:
        xin = ops.split_bf16(x.reshape(-1, Cin).to(DEV), ops.SPLIT_TERMS[mode])
        ref_in, ref_w = x.double(), w.double()
        tol = 3e-5
    acc = _ref_conv(ref_in, ref_w, b.double(), kind)
    ref = O.aptx((acc + res.double()).masked_fill(m, 0.0), 0.9, 0.6)          # ResidualBlock1D tail (attentions.py:545-549)
    pc = ops.pack_conv(w if mode != "bf16" else wq.float(), b, kind, split=mode).to(DEV)
    outs = {}
    for pair in (True, False):
        o32 = torch.empty(N, H, 1, Cout, dtype=torch.float32, device=DEV)
        ops.conv_gemm(xin, pc, N, H, 1, row_mask=mask.to(DEV), mask_pre=True, act=True, beta=0.9, gamma=0.6,
                      fast_tanh=False, res=res.to(DEV), res_mode=1, out_f32=o32, pair=pair)
        outs[pair] = o32.cpu()
    scale = max(1.0, ref.abs().max().item())
    assert (outs[True].double() - ref).abs().max().item() < tol * scale
    assert (outs[True] - outs[False]).abs().max().item() < tol * scale
    if mode == "bf16":               # lean bf16 epilogue with a bf16 residual, run twice (ring state across launches)
        rb = res.to(torch.bfloat16)
        ref2 = O.aptx((acc + rb.double()).masked_fill(m, 0.0), 0.9, 0.6)
        for _ in range(2):
            ob = torch.zeros(N, H, 1, Cout, dtype=torch.bfloat16, device=DEV)
            ops.conv_gemm(xin, pc, N, H, 1, row_mask=mask.to(DEV), mask_pre=True, act=True, beta=0.9, gamma=0.6,
                          res=rb.to(DEV), res_mode=1, out_bf16=ob, pair=True)
            assert (ob.cpu().double() - ref2).abs().max().item() < 2e-2 * max(1.0, ref2.abs().max().item())


@pytest.mark.parametrize("N,H,W,Cin,Cout,pair", [
    (2, 64, 24, 64, 64, True),           # narrow layer, msub 4 on the CTA-pair kernel, three product segments
    (1, 64, 144, 192, 64, True),         # ups.2.conv1 shape: three 64-channel chunks per segment
    (1, 32, 16, 128, 256, True),
    (2, 24, 20, 64, 96, False),          # H < 32: generic tap loop with segments
])
def test_conv3x3_f16x2_on_pair_kernel_is_fp32_grade(N, H, W, Cin, Cout, pair):
    """fp32-grade decoder mode: a 3x3 convolution as three fp16 products of 2-term operand splits on conv_pair_kernel
    (segment-major K loop over the halo), against float64; error within 4x the reference's own fp32 error."""
    x = _rand(N, H, W, Cin, seed=60) * 2.0
    w = _rand(Cout, Cin, 3, 3, seed=61) / (9 * Cin) ** 0.5
    b = _rand(Cout, seed=62)
    ref = _ref_conv(x.double(), w.double(), b.double(), "conv2d3")
    ref32 = _ref_conv(x, w, b, "conv2d3")
    err32 = (ref32.double() - ref).abs().max().item()
    pc = ops.pack_conv(w, b, "conv2d3", split="f16x2").to(DEV)
    xs = ops.split_bf16(x.reshape(-1, Cin).to(DEV), 2)
    out = torch.empty(N, H, W, Cout, dtype=torch.float32, device=DEV)
    osp = torch.empty(N, H, W, 2 * Cout, dtype=torch.float16, device=DEV)
    mask = (torch.arange(N * H) % 7 == 3).to(torch.uint8).to(DEV)
    ops.conv_gemm(xs, pc, N, H, W, out_f32=out, out_split=osp, pair=pair if pair else None, row_mask=mask, mask_post=True)
    ref = ref.masked_fill(mask.cpu().bool().reshape(N, H, 1, 1), 0.0)
    err = (out.cpu().double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err < max(4 * err32, 2e-5 * scale), (err, err32, scale)
    s2 = osp.cpu().float().reshape(N, H, W, 2, Cout).sum(dim=3)
    assert (s2 - out.cpu()).abs().max().item() < 1e-6 * max(1.0, scale) + 2e-7
    if pair:       # the generic tap loop gives the same numbers up to accumulation order
        out2 = torch.empty_like(out)
        ops.conv_gemm(xs, pc, N, H, W, out_f32=out2, pair=False, halo=False, row_mask=mask, mask_post=True)
        assert (out2 - out).abs().max().item() < 1e-5 * max(1.0, scale)


def test_split_pool_upcat_stem_for_fp32_grade_decoder():
    """mq_avgpool_mask_split / mq_upcat_mask_split / mq_refiner_stem_split against the reference ops in float64."""
    B, T, depth, Fw, Cc = 2, 21, 3, 20, 16
    lengths = torch.tensor([21, 13])
    mask = O.sequence_mask(T, lengths)
    T8, down, up = ops.refiner_masks(mask.to(torch.uint8).to(DEV), B, T, depth, DEV)
    x = _rand(B, T8, Fw, Cc, seed=70) * 3.0
    xs = ops.split_bf16(x.reshape(-1, Cc).to(DEV), 2).reshape(B, T8, Fw, 2 * Cc)

    def join(t, c):
        return t.cpu().float().reshape(*t.shape[:-1], 2, c).sum(dim=-2)

    assert (join(xs, Cc) - x).abs().max().item() < 2e-6
    m1 = down[1].cpu().bool().reshape(B, T8 // 2, 1, 1)
    y = ops.avgpool_mask_split(xs, down[1], B, T8, Fw, Cc)
    ref = (0.5 * (x[:, 0::2] + x[:, 1::2])).masked_fill(m1, 0.0)
    assert (join(y, Cc) - ref).abs().max().item() < 4e-6
    lo = _rand(B, T8 // 2, Fw, 2 * Cc, seed=71)
    los = ops.split_bf16(lo.reshape(-1, 2 * Cc).to(DEV), 2).reshape(B, T8 // 2, Fw, 4 * Cc)
    u = ops.upcat_mask_split(los, xs, up[0], B, T8, Fw, 2 * Cc, Cc)
    m0 = up[0].cpu().bool().reshape(B, T8, 1, 1)
    ref = torch.cat([lo.repeat_interleave(2, dim=1), x], dim=-1).masked_fill(m0, 0.0)
    assert tuple(u.shape) == (B, T8, Fw, 2 * 3 * Cc)
    assert (join(u, 3 * Cc) - ref).abs().max().item() < 4e-6
    # exact copy semantics: term 0 of the concat is [lo_h0 | x_h0]
    assert torch.equal(u[..., :2 * Cc][~m0.squeeze(-1).squeeze(-1).to(DEV)],
                       los[..., :2 * Cc].repeat_interleave(2, dim=1)[~m0.squeeze(-1).squeeze(-1).to(DEV)])
    r = _rand(B, T, Fw, seed=72)
    w1, b1 = _rand(Cc, 1, 3, 3, seed=73) * 0.3, _rand(Cc, seed=74) * 0.1
    md = down[0].cpu().bool().reshape(B, 1, T8, 1)
    img = torch.cat([r, torch.zeros(B, T8 - T, Fw)], 1).unsqueeze(1).masked_fill(md, 0.0)
    ref = O.aptx(F.conv2d(img.double(), w1.double(), b1.double(), padding=1), 1, 0.5).permute(0, 2, 3, 1)
    s = ops.refiner_stem_split(r.to(DEV), mask.to(torch.uint8).to(DEV), B, T, T8, Fw, Cc,
                               w1.reshape(Cc, 9).contiguous().to(DEV), b1.to(DEV))
    assert (join(s, Cc).double() - ref).abs().max().item() < 2e-6 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("n,K,D", [(3000, 1000, 20), (4097, 512, 4), (2048, 8192, 5)])
def test_vq_nearest_folded_and_unfolded_forms_agree(n, K, D):
    """||c||^2 inside the GEMM (augmented K columns, the default for d <= 61) and ||c||^2 added in the epilogue give the
    same argmin as the float64 reference and the same distances to fp32 accuracy."""
    z = _rand(n, D, seed=51) * 2.0
    cb = _rand(K, D, seed=52) * 3.0
    ref_idx, gap, d64 = _vq_ref(z, cb)
    scale = float(d64.min(dim=1).values.mean()) + 1.0
    out = {}
    for fold in (True, False):
        pc = ops.pack_codebook(cb, "f16x2", fold=fold).to(DEV)
        assert pc.fold == fold
        idx, dist = ops.vq_nearest(z.to(DEV), pc, want_codes=False, want_dist=True)
        neq = idx.cpu() != ref_idx
        assert int((neq & (gap > 1e-5 * scale)).sum()) == 0, (fold, int(neq.sum()))
        out[fold] = (idx.cpu(), dist.cpu())
    assert float((out[True][0] == out[False][0]).float().mean()) > 0.9995
    assert (out[True][1] - out[False][1]).abs().max().item() < 2e-5 * scale
    # bf16 form: folded ||c||^2 keeps 24 bits (three bf16 terms), so its agreement matches the unfolded bf16 form
    ag = []
    for fold in (True, False):
        ib = ops.vq_nearest(z.to(DEV), ops.pack_codebook(cb, "bf16", fold=fold).to(DEV), want_codes=False).cpu()
        ag.append(float((ib == ref_idx).float().mean()))
    print("bf16 agreement folded / unfolded", ag)
    assert abs(ag[0] - ag[1]) < 0.02
