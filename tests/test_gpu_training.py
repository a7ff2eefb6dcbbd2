"""Gradient kernels of the training step on the GPU, through the C ABI.

The weight-gradient kernel (mq_conv_wgrad) and the data gradient (mq_conv_gemm on the mirrored,
transposed weight) are compared with float64 autograd of the same reference op (F.conv2d / F.conv1d /
F.linear, the calls the reference's loss.backward() differentiates) on bf16-rounded operands, so only
the fp32 accumulation order differs.  Tolerances are written beside each assert.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from mqgan_b200 import ops  # noqa: E402

DEV = "cuda"


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def _ref_conv(x, w, kind):
    """x (N,H,W,C) float64 channel-last -> (N,H,W,Cout) float64 (no bias)."""
    N, H, W, Cc = x.shape
    if kind == "linear":
        return F.linear(x, w)
    if kind in ("same1d", "causal1d"):
        k = w.shape[2]
        xi = x.reshape(N, H, Cc).permute(0, 2, 1)
        if kind == "causal1d":
            y = F.conv1d(F.pad(xi, (k - 1, 0)), w)
        else:
            y = F.conv1d(xi, w, padding=(k - 1) // 2)
        return y.permute(0, 2, 1).reshape(N, H, 1, -1)
    if kind == "conv2d3":
        return F.conv2d(x.permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1)
    raise ValueError(kind)


def _ref_grads(x, w, dy, kind):
    x = x.double().requires_grad_(True)
    w = w.double().requires_grad_(True)
    y = _ref_conv(x, w, kind)
    y.backward(dy.double())
    return x.grad, w.grad


GRAD_CASES = [
    # kind, N, H, W, Cin, Cout, wshape-tail
    ("linear", 1, 256, 1, 128, 64, ()),              # one A atom (cout 64), two B atoms
    ("linear", 3, 77, 1, 512, 144, ()),              # ragged rows (77 = 64 + 13), cout 144 -> two co tiles
    ("same1d", 2, 300, 1, 64, 96, (3,)),
    ("same1d", 1, 130, 1, 192, 256, (5,)),           # bn = 192
    ("causal1d", 2, 200, 1, 768, 512, (7,)),         # three ci tiles, four co tiles, 7 taps
    ("conv2d3", 2, 24, 144, 64, 128, (3, 3)),        # refiner-like image, 8x8 pixel boxes
    ("conv2d3", 1, 7, 36, 96, 192, (3, 3)),          # ragged H and W, Cin not a multiple of 64
    ("conv2d3", 1, 16, 144, 192, 64, (3, 3)),
    ("conv2d3", 1, 32, 20, 384, 256, (3, 3)),        # bn = 192 x 2
]


@pytest.mark.parametrize("kind,N,H,W,Cin,Cout,tail", GRAD_CASES)
def test_conv_wgrad_matches_autograd(kind, N, H, W, Cin, Cout, tail):
    x = _rand(N, H, W, Cin, seed=21).to(torch.bfloat16)
    w = (_rand(Cout, Cin, *tail, seed=22) / (Cin * max(1, int(np.prod(tail)))) ** 0.5).to(torch.bfloat16)
    dy = _rand(N, H, W, Cout, seed=23).to(torch.bfloat16)
    _, gw = _ref_grads(x, w, dy, kind)
    dh, dw = ops.conv_taps(kind, w.shape)
    ref = gw.reshape(Cout, Cin, -1).permute(2, 0, 1)                      # (taps, cout, cin)
    scale = max(1.0, ref.abs().max().item())
    for split in (None, 1, 3):
        out = ops.conv_wgrad(dy.to(DEV), x.to(DEV), N, H, W, Cout, Cin, dh, dw, split=split)
        torch.cuda.synchronize()
        assert out.shape == ref.shape
        err = (out.cpu().double() - ref).abs().max().item()
        # bf16 operands are exact on both sides; fp32 accumulation over N*H*W pixels
        assert err < 3e-4 * scale, (split, err, scale)


@pytest.mark.parametrize("kind,N,H,W,Cin,Cout,tail", GRAD_CASES)
def test_conv_dgrad_is_forward_kernel_on_mirrored_weight(kind, N, H, W, Cin, Cout, tail):
    x = _rand(N, H, W, Cin, seed=31).to(torch.bfloat16)
    w = (_rand(Cout, Cin, *tail, seed=32) / (Cin * max(1, int(np.prod(tail)))) ** 0.5).to(torch.bfloat16)
    dy = _rand(N, H, W, Cout, seed=33).to(torch.bfloat16)
    gx, _ = _ref_grads(x, w, dy, kind)
    wd, kd = ops.dgrad_weight(w.float().to(DEV), kind)
    pc = ops.pack_conv(wd, None, kd, on_device=True)
    assert pc.wpack.is_cuda
    out = torch.empty(N, H, W, Cin, dtype=torch.float32, device=DEV)
    ops.conv_gemm(dy.to(DEV), pc, N, H, W, out_f32=out)
    torch.cuda.synchronize()
    err = (out.cpu().double() - gx).abs().max().item()
    assert err < 2e-4 * max(1.0, gx.abs().max().item()), err


def test_conv_wgrad_large_k_split_is_deterministic():
    """Config-5-sized pixel count (16 x 256 frames at the refiner's 1/8 resolution x 144 bins) on a wide
    layer: two runs are bit-identical (no atomics), and linear in dy."""
    N, H, W, Cin, Cout = 16, 32, 144, 256, 256
    x = _rand(N, H, W, Cin, seed=41).to(torch.bfloat16).to(DEV)
    dy = _rand(N, H, W, Cout, seed=42).to(torch.bfloat16).to(DEV)
    dh, dw = ops.conv_taps("conv2d3", (Cout, Cin, 3, 3))
    a = ops.conv_wgrad(dy, x, N, H, W, Cout, Cin, dh, dw)
    b = ops.conv_wgrad(dy, x, N, H, W, Cout, Cin, dh, dw)
    assert torch.equal(a, b)
    c = ops.conv_wgrad((dy.float() * 2).to(torch.bfloat16), x, N, H, W, Cout, Cin, dh, dw)
    assert torch.equal(c, 2 * a)                                          # power-of-two scaling is exact
    # centre tap against a plain matmul over pixels
    ref = dy.float().reshape(-1, Cout).t().double() @ x.float().reshape(-1, Cin).double()
    err = (a[4].double() - ref).abs().max().item()
    assert err < 3e-4 * max(1.0, ref.abs().max().item()), err


# ----------------------------------------------------------------------------
# ConvBlock2D point-wise stage, forward and backward (mq_cb2d_point_forward / mq_cb2d_backward)
# ----------------------------------------------------------------------------
@pytest.mark.parametrize("B,T,C", [(2, 37, 64), (1, 300, 96), (3, 8, 512)])
def test_cb2d_point_forward_backward_match_float64_autograd(B, T, C):
    s = _rand(B, T, C, seed=51) * 1.5
    wpw, bpw, wout = _rand(C, seed=52), _rand(C, seed=53) * 0.3, _rand(C, seed=54) / C ** 0.5
    bout = torch.tensor([0.37])
    dy = _rand(B, T, C, seed=55)
    mask = torch.zeros(B, T, dtype=torch.bool)
    mask[0, T // 2:] = True
    s = s.masked_fill(mask.unsqueeze(-1), 0.0)
    # float64 definition (preencoder.py:288-295 on a masked s)
    p = [t.double().requires_grad_(True) for t in (s, wpw, bpw, wout, bout)]
    u = (p[0].unsqueeze(-1) * p[1] + p[2]).masked_fill(mask[:, :, None, None], 0.0)
    y_ref = (((1 + torch.tanh(u)) * 0.5 * u) * p[3]).sum(-1) + p[4]
    y_ref.backward(dy.double())
    m8 = mask.to(torch.uint8).to(DEV)
    y = ops.cb2d_point_forward(s.to(DEV), wpw.to(DEV), bpw.to(DEV), wout.to(DEV), bout.to(DEV), m8)
    ds, dwpw, dbpw, dwout, dbout = ops.cb2d_point_backward(s.to(DEV), dy.to(DEV), wpw.to(DEV), bpw.to(DEV), wout.to(DEV), m8)
    torch.cuda.synchronize()

    def close(a, b, tol):
        return float((a.cpu().double() - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))

    assert close(y, y_ref.detach(), 5e-6)            # fp32 sums of C terms, tanh to ~1e-7
    assert close(ds, p[0].grad, 5e-6)
    assert close(dwpw, p[1].grad, 2e-5)              # fp32 sums over B*T*C pixels
    assert close(dbpw, p[2].grad, 2e-5)
    assert close(dwout, p[3].grad, 2e-5)
    assert close(dbout, p[4].grad, 2e-5)
    # tanh.approx mode (the training default: the kernels are MUFU-bound): ~2^-11 relative on each tanh
    yf = ops.cb2d_point_forward(s.to(DEV), wpw.to(DEV), bpw.to(DEV), wout.to(DEV), bout.to(DEV), m8, fast_tanh=True)
    gf = ops.cb2d_point_backward(s.to(DEV), dy.to(DEV), wpw.to(DEV), bpw.to(DEV), wout.to(DEV), m8, fast_tanh=True)
    assert close(yf, y_ref.detach(), 2e-3)
    for got, ref in zip(gf, (p[0].grad, p[1].grad, p[2].grad, p[3].grad, p[4].grad)):
        assert close(got, ref, 4e-3)


# ----------------------------------------------------------------------------
# the whole training iteration against the reference's own Trainer step (tests/golden/train_tiny.npz)
# ----------------------------------------------------------------------------
def _tiny_train_step(native_cb2d=True, name="train_tiny", **kw):
    import os
    from mqgan_b200 import spec as S
    from mqgan_b200 import training as TR
    from mqgan_b200.synth import synth_disc_state_dict, synth_state_dict
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    cfg, pdc, mbc = (getattr(S, str(n)) for n in fx["configs"])
    seed = int(fx["seed"])
    g_sd = synth_state_dict(cfg, seed=seed)
    g_sd["q_in_proj.weight"] = torch.from_numpy(fx["qin_w"]).clone()
    g_sd["q_in_proj.bias"] = torch.from_numpy(fx["qin_b"]).clone()
    pd_sd = synth_disc_state_dict(S.patch_disc_param_spec(pdc), seed=seed)
    mb_sd = synth_disc_state_dict(S.multibin_param_spec(mbc), seed=seed + 1)
    ts = TR.TrainStep(cfg, pdc, mbc, g_sd, pd_sd, mb_sd, dict(S.TINY_TRAIN), DEV, native_cb2d=native_cb2d, **kw)
    return fx, cfg, ts


def _tiny_batch(step, B, T, n_mels):
    from mqgan_b200.synth import synth_lengths, synth_mels
    real = synth_mels(B, T, n_mels, seed=40 + step)
    lens = synth_lengths(B, T, seed=40 + step, ragged=True)
    return real.masked_fill((torch.arange(T)[None, :] >= lens[:, None]).unsqueeze(-1), 0.0), lens


@pytest.mark.parametrize("name", ["train_tiny", "train_tiny_m"])
def test_train_step_matches_reference_trainer_two_iterations(name):
    """Losses, reconstructions, gradients (as the step leaves them: clipped), updated weights, spectral-norm
    vectors and LeCam anchors after two consecutive iterations (the second with feature matching) against the
    UNMODIFIED reference run on the CPU in fp32.  The generator's convolutions run with bf16 operands here
    (the reference's own CUDA training precision, train.py:523), everything else in fp32, TF32 off.
    Measured on B200 (tools/train_parity.py, profiles/train_parity_r01.json) next to each bound."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    fx, cfg, ts = _tiny_train_step(name=name)
    B, T = int(fx["B"]), int(fx["T"])
    g_keys = [str(k) for k in fx["g_keys"]]
    names = ["loss_d", "loss_g_total", "loss_recon_pre", "loss_recon_post", "loss_gan", "loss_fm"]
    for step in (1, 2):
        real, lens = _tiny_batch(step, B, T, cfg.mel_channels)
        o = ts.step(real, lens, gan=True, use_fm=step == 2)
        pre = f"s{step}_"
        got = np.array([float(o[n]) for n in names])
        np.testing.assert_allclose(got, fx[pre + "losses"], rtol=1e-3, atol=1e-6)                  # measured <= 1.1e-5
        rp, rq = ts.last_recon
        assert float((rp.cpu() - torch.from_numpy(fx[pre + "recon_pre"])).abs().max()) < 5e-3      # measured 1.9e-4 (|mel| ~ 10)
        assert float((rq.cpu() - torch.from_numpy(fx[pre + "recon_post"])).abs().max()) < 5e-3     # measured 5.2e-4
        gn = np.array([float(ts.g[k].grad.norm()) for k in g_keys])
        refn = fx[pre + "g_grad_norms"]
        for k, r in zip(g_keys, refn):
            if r < 0:                                   # hidden_proj: no gradient in the reference (preencoder.py:411-413)
                assert float(ts.g[k].grad.abs().max()) == 0.0, k
        has = refn > 1e-4 * refn.max()                  # below that the reference's own gradient is float noise
        rel = np.abs(gn[has] - refn[has]) / refn[has]
        assert np.median(rel) < 1e-2, np.median(rel)    # measured 1.5e-3
        # worst case measured 0.13: encoder_blocks.2.relu.beta, a scalar whose gradient is a cancelling sum over every
        # activation of the block, in the second iteration
        assert rel.max() < 0.3, (rel.max(), g_keys[int(np.flatnonzero(has)[int(rel.argmax())])])
        for name in fx.files:
            if name.startswith(pre + "gg:"):
                k = name[len(pre) + 3:]
                b = torch.from_numpy(fx[name]).double().reshape(-1)
                if float(b.norm()) <= 1e-4 * refn.max():
                    continue
                a = ts.g[k].grad.detach().cpu().double().reshape(-1)
                # bf16 operands in fwd / dgrad / wgrad: measured 0.1 - 4.3 % on tensors; 13 % on the scalar relu.beta
                assert float((a - b).norm() / b.norm()) < (0.3 if b.numel() == 1 else 8e-2), k
        ps = np.array([float(ts.g[k].detach().double().sum()) for k in g_keys])
        assert np.abs(ps - fx[pre + "g_param_sums"]).max() < 2e-2                                  # measured 2.3e-3 (Adam: |delta| = lr per element)
        assert float((ts.pd["convs.1.weight_u"].cpu() - torch.from_numpy(fx[pre + "d_u0"])).abs().max()) < 1e-5
        np.testing.assert_allclose(ts.lecam.ema.cpu().numpy(), fx[pre + "lecam"], rtol=1e-3, atol=1e-5)
        d_now = {**{"pd:" + k: v for k, v in ts.pd.items()}, **{"mb:" + k: v for k, v in ts.mb.items()}}
        dps = np.array([float(d_now[str(k)].detach().double().sum()) for k in fx["d_keys"]])
        assert np.abs(dps - fx[pre + "d_param_sums"]).max() < 1e-3                                 # measured 3.7e-5


def test_train_step_needs_cuda_and_zero_dropout():
    from mqgan_b200 import spec as S
    from mqgan_b200 import training as TR
    with pytest.raises(RuntimeError):
        TR.TrainStep(S.TINY, S.TINY_PATCH_D, S.TINY_MULTIBIN_D, {}, {}, {}, dict(S.TINY_TRAIN), "cpu")
    with pytest.raises(NotImplementedError):
        TR.TrainStep(S.TINY, S.TINY_PATCH_D, S.TINY_MULTIBIN_D, {}, {}, {}, dict(S.TINY_TRAIN), DEV, dropout_p=0.1)


@pytest.mark.parametrize("C,residual", [(64, True), (128, False)])
def test_act_forward_backward_match_float64(C, residual):
    """mq_act_forward / mq_act_backward (APTx [+ residual] + row mask, bf16 in/out) against float64 autograd."""
    N, H, W = 2, 9, 20
    u = _rand(N, H, W, C, seed=61) * 2
    res = _rand(N, H, W, C, seed=62).to(torch.bfloat16)
    dy = _rand(N, H, W, C, seed=63).to(torch.bfloat16)
    mask = torch.zeros(N, H, dtype=torch.bool)
    mask[1, 5:] = True
    ud = u.double().requires_grad_(True)
    rd = res.double().requires_grad_(True)
    y_ref = ((1 + torch.tanh(ud)) * 0.5 * ud + (rd if residual else 0.0)).masked_fill(mask[:, :, None, None], 0.0)
    y_ref.backward(dy.double())
    m8 = mask.to(torch.uint8).to(DEV)
    y = ops.act_forward(u.to(DEV), res.to(DEV) if residual else None, m8, W)
    du, dres, db = ops.act_backward(dy.to(DEV), u.to(DEV), m8, W, want_res=residual, want_bias=True)
    du_only, _ = ops.act_backward(dy.to(DEV), u.to(DEV), m8, W)
    torch.cuda.synchronize()
    assert torch.equal(du, du_only)
    ref_db = ud.grad.sum(dim=(0, 1, 2))
    assert float((db.cpu().double() - ref_db).abs().max()) <= 1e-5 * max(1.0, float(ref_db.abs().max()))   # fp32 column sums
    assert y.dtype == torch.bfloat16 and du.dtype == torch.bfloat16
    # outputs are rounded to bf16 once: half an ulp = 2^-9 relative
    assert float((y.cpu().double() - y_ref.detach()).abs().max()) <= 2.0 ** -8 * float(y_ref.abs().max())
    assert float((du.cpu().double() - ud.grad).abs().max()) <= 2.0 ** -8 * float(ud.grad.abs().max())
    if residual:
        assert torch.equal(dres.cpu().double(), rd.grad)          # dy passed through (exact) or zeroed
    else:
        assert dres is None


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_leaky_mask_forward_backward(dtype):
    """mq_leaky_mask_forward / backward on a channels_last feature map against the torch definition."""
    B, Cc, H, W = 2, 32, 5, 7                      # 256 % (C / 8) == 0: the fused bias gradient is available
    y = _rand(B, Cc, H, W, seed=71).to(dtype).to(DEV).contiguous(memory_format=torch.channels_last)
    dout = _rand(B, Cc, H, W, seed=72).to(torch.bfloat16).to(DEV).contiguous(memory_format=torch.channels_last)
    mask = torch.zeros(B, 1, H, W, dtype=torch.bool, device=DEV)
    mask[1, :, :, 4:] = True
    m8 = mask.reshape(B, H, W).to(torch.uint8).contiguous()
    bias = _rand(Cc, seed=73).to(DEV)
    for b in (None, bias):
        yb = y.float() if b is None else y.float() + b.reshape(1, -1, 1, 1)
        out = ops.leaky_mask_forward(y, m8, 0.2, b)
        du, db = ops.leaky_mask_backward(dout, y, m8, 0.2, b, want_bias=True)
        ref = F.leaky_relu(yb, 0.2).masked_fill(mask, 0.0)
        assert out.dtype == torch.bfloat16 and out.is_contiguous(memory_format=torch.channels_last)
        assert torch.equal(out, ref.to(torch.bfloat16))
        ref_du = (dout.float() * torch.where(yb > 0, 1.0, 0.2)).masked_fill(mask, 0.0)
        assert torch.equal(du, ref_du.to(torch.bfloat16))
        assert torch.equal(du, ops.leaky_mask_backward(dout, y, m8, 0.2, b))
        ref_db = ref_du.double().sum(dim=(0, 2, 3))
        assert float((db.double() - ref_db).abs().max()) <= 1e-5 * max(1.0, float(ref_db.abs().max()))


def test_discriminator_bf16_mode_tracks_fp32_mode():
    """The reference-precision (bf16 autocast) discriminator path - channels_last bf16 feature maps, fused
    LeakyReLU + mask passes - against the fp32 path on the same weights: logits and input gradients."""
    from mqgan_b200 import spec as S
    from mqgan_b200 import training as TR
    from mqgan_b200.synth import synth_disc_state_dict, synth_mels
    dc = S.PatchDiscConfig(32, (16, 32, 64), ((5, 5), (5, 5), (3, 3), (3, 3)), ((1, 2), (2, 2), (2, 1), (1, 1)))
    sd = {k: v.to(DEV) for k, v in synth_disc_state_dict(S.patch_disc_param_spec(dc), seed=9).items()}
    lens = torch.tensor([40, 23], device=DEV)
    with torch.no_grad():                          # settle the spectral-norm vectors (random u, v give a near-zero sigma)
        for _ in range(20):
            TR.patch_discriminator(sd, dc, synth_mels(2, 40, 32, seed=3).to(DEV), lens, training=True)
    outs, grads = [], []
    for fast, native in ((False, False), (True, False), (True, True)):
        x = synth_mels(2, 40, 32, seed=3).to(DEV).requires_grad_(True)
        logits, valid, feats = TR.patch_discriminator(sd, dc, x, lens, training=False, autocast_bf16=fast, native_conv=native)
        (logits * valid).pow(2).sum().backward()
        outs.append(logits.detach())
        grads.append(x.grad.detach())
        assert len(feats) == 1
    for k in (1, 2):                               # cuDNN bf16 path, then every convolution on the tcgen05 kernels
        e_out = float((outs[0] - outs[k]).abs().max()) / float(outs[0].abs().max())
        e_grad = float((grads[0] - grads[k]).norm() / grads[0].norm())
        assert e_out < 3e-2, (k, e_out)            # bf16 operands and bf16 feature maps through four layers
        assert e_grad < 0.15, (k, e_grad)
    # weight gradients of the native path against the fp32 path
    gw = []
    for fast, native in ((False, False), (True, True)):
        sdg = {k: (v.clone().requires_grad_(True) if not S.is_disc_buffer(k) else v.clone()) for k, v in sd.items()}
        x = synth_mels(2, 40, 32, seed=3).to(DEV)
        logits, valid, _ = TR.patch_discriminator(sdg, dc, x, lens, training=False, autocast_bf16=fast, native_conv=native)
        (logits * valid).pow(2).sum().backward()
        gw.append({k: v.grad for k, v in sdg.items() if v.requires_grad})
    for k in gw[0]:
        ref = gw[0][k]
        assert float((gw[1][k] - ref).norm()) <= 0.15 * float(ref.norm()) + 1e-6 * float(ref.abs().max() + 1), k


def test_graph_replayed_iterations_equal_eager_iterations():
    """TrainStep.capture() + step_graphed(): the whole iteration replayed from one CUDA graph trains exactly like
    the eager launches (same kernels, same order), including the warm-up learning-rate schedule fed through a
    device tensor."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    fx, cfg, eager = _tiny_train_step()
    _, _, graphed = _tiny_train_step()
    B, T = int(fx["B"]), int(fx["T"])
    batches = [_tiny_batch(s, B, T, cfg.mel_channels) for s in (1, 2, 3)]
    for _ in range(3):
        eager.step(*batches[0])
    graphed.capture(batches[0][0], batches[0][1], warmup=3)
    assert graphed.g_steps == eager.g_steps == 3
    for real, lens in batches[1:]:
        a = eager.step(real, lens)
        b = graphed.step_graphed(real, lens)
        for k in a:
            assert abs(float(a[k]) - float(b[k])) <= 1e-5 * max(1.0, abs(float(a[k]))), k
    assert graphed.g_steps == eager.g_steps == 5
    assert float(graphed.lr_g) == float(eager.lr_g) > 0
    for k in eager.g:
        d = float((eager.g[k].detach() - graphed.g[k].detach()).abs().max())
        assert d <= 1e-6 + 1e-4 * float(eager.g[k].detach().abs().max()), (k, d)
    with pytest.raises(KeyError):
        graphed.step_graphed(batches[0][0][:2], batches[0][1][:2])


def test_dgrad_wider_than_one_launch_is_sliced():
    """A data gradient with more than 1024 output channels (hifimusic's first up-block reads 768 + 384 = 1152 channels)
    is produced in channel slices; against float64 autograd."""
    from mqgan_b200 import training as TR
    N, H, W, Cin, Cout = 1, 16, 24, 1152, 64
    x = _rand(N, H, W, Cin, seed=81).to(torch.bfloat16)
    w = (_rand(Cout, Cin, 3, 3, seed=82) / (9 * Cin) ** 0.5).to(torch.bfloat16)
    dy = _rand(N, H, W, Cout, seed=83).to(torch.bfloat16)
    gx, _ = _ref_grads(x, w, dy, "conv2d3")
    for dt in (torch.float32, torch.bfloat16):
        out = torch.empty(N, H, W, Cin, dtype=dt, device=DEV)
        TR._dgrad_launch(dy.to(DEV), w.float().to(DEV), "conv2d3", N, H, W, out, "t")
        tol = 2e-4 if dt == torch.float32 else 2.0 ** -7
        assert float((out.cpu().double() - gx).abs().max()) < tol * max(1.0, float(gx.abs().max()))


def test_preencoder_module_trains_through_forward():
    """The drop-in nn.Module: in train mode with autograd on, PreEncoder.forward is the differentiable training
    forward over the module's own parameters (what the reference's Trainer calls, train.py:524), so a plain torch
    optimiser loop works; in eval mode the same module keeps serving encode / decode from the inference engine and
    sees the updated weights."""
    from mqgan_b200 import spec as S
    from mqgan_b200 import training as TR
    from mqgan_b200.preencoder import PreEncoder
    from mqgan_b200.synth import synth_state_dict
    fx, cfg, ts = _tiny_train_step()
    sd = synth_state_dict(cfg, seed=int(fx["seed"]))
    sd["q_in_proj.weight"] = torch.from_numpy(fx["qin_w"]).clone()
    sd["q_in_proj.bias"] = torch.from_numpy(fx["qin_b"]).clone()
    model = PreEncoder(cfg.mel_channels, list(cfg.channels), list(cfg.kernel_sizes), fsq_levels=list(cfg.fsq_levels), dropout=0.0,
                       refiner_base_channels=cfg.refiner_base_channels, refiner_depth=cfg.refiner_depth)
    model.load_state_dict(sd, strict=True)
    model.to(DEV)
    real, lens = _tiny_batch(1, int(fx["B"]), int(fx["T"]), cfg.mel_channels)
    real, lens = real.to(DEV), lens.to(DEV)
    model.eval()
    idx0 = model.encode(real, None)
    model.train()
    x_recon, x_post = model(real, lens)
    assert x_recon.requires_grad and x_post.requires_grad
    # same numbers as the functional training forward on the same parameters, and as the reference (fixture s1)
    r2, p2 = TR.generator_forward(dict(model.named_parameters()), cfg, real, lens)
    assert torch.equal(x_recon, r2) and torch.equal(x_post, p2)
    assert float((x_post.detach().cpu() - torch.from_numpy(fx["s1_recon_post"])).abs().max()) < 5e-3
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    loss = TR.masked_mel_loss(x_post, real, lens, 1) + TR.masked_mel_loss(x_recon, real, lens, 1)
    loss.backward()
    got = {k for k, p in model.named_parameters() if p.grad is not None}
    assert "refiner.mid.conv1.parametrizations.weight.original1" in got and "proj.weight" in got
    assert "hidden_proj.weight" not in got                         # detached refiner input (preencoder.py:411-413)
    opt.step()
    with torch.no_grad():                                          # inference forward of a training-mode module: engine path
        a, b = model(real, lens)
    assert not a.requires_grad and a.shape == real.shape
    model.eval()
    out = model.decode(model.encode(real, None), None)             # the engine re-packs the updated weights
    assert out.shape == real.shape and torch.isfinite(out).all()
    assert model.encode(real, None).shape == idx0.shape


def test_reference_style_training_loop_on_dropin_modules():
    """The reference's own loop structure (Trainer._train_discriminator / _train_generator, train.py:380-501) written
    against the drop-in classes - mqgan_b200.PreEncoder as the generator, mqgan_b200.discriminators, mqgan_b200.losses,
    torch.optim.Adam / LambdaLR / clip_grad_norm_ exactly as train.py uses them - reproduces the reference's logged
    losses for two iterations (tests/golden/train_tiny.npz)."""
    from mqgan_b200 import spec as S
    from mqgan_b200.discriminators import MelSpectrogramPatchDiscriminator2D, MultiBinDiscriminator
    from mqgan_b200.losses import LSGANLoss, MaskedMelLoss
    from mqgan_b200.preencoder import PreEncoder
    from mqgan_b200.synth import synth_disc_state_dict, synth_state_dict
    from mqgan_b200.training import masked_mae
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import os
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_tiny.npz"))
    cfg, pdc, mbc, t = S.TINY, S.TINY_PATCH_D, S.TINY_MULTIBIN_D, S.TINY_TRAIN
    seed = int(fx["seed"])
    sd = synth_state_dict(cfg, seed=seed)
    sd["q_in_proj.weight"], sd["q_in_proj.bias"] = torch.from_numpy(fx["qin_w"]).clone(), torch.from_numpy(fx["qin_b"]).clone()
    generator = PreEncoder(cfg.mel_channels, list(cfg.channels), list(cfg.kernel_sizes), fsq_levels=list(cfg.fsq_levels),
                           dropout=0.0, refiner_base_channels=cfg.refiner_base_channels, refiner_depth=cfg.refiner_depth)
    generator.load_state_dict(sd, strict=True)
    patch_d = MelSpectrogramPatchDiscriminator2D(pdc.mel_channels, list(pdc.hidden_channels), [k[0] for k in pdc.kernels],
                                                 stride=[list(s) for s in pdc.strides], fast=False)
    patch_d.load_state_dict(synth_disc_state_dict(S.patch_disc_param_spec(pdc), seed=seed), strict=True)
    multibin_d = MultiBinDiscriminator(mbc.mel_channels, mbc.n_bins, list(mbc.hidden_channels), list(mbc.kernel_sizes),
                                       mbc.n_no_strides, fast=False)
    multibin_d.load_state_dict(synth_disc_state_dict(S.multibin_param_spec(mbc), seed=seed + 1), strict=True)
    for m in (generator, patch_d, multibin_d):
        m.to(DEV)
    gan_loss = LSGANLoss().to(DEV)
    recon_all, recon_group = MaskedMelLoss("mse"), MaskedMelLoss("mse", group_size=16)
    opt_g = torch.optim.Adam(generator.parameters(), lr=t["lr"], betas=(t["beta1"], t["beta2"]))
    d_params = list(patch_d.parameters()) + list(multibin_d.parameters())
    opt_d = torch.optim.Adam(d_params, lr=t["lr"] * t["lr_d_factor"], betas=(t["d_beta1"], t["d_beta2"]))
    sched_g = torch.optim.lr_scheduler.LambdaLR(opt_g, lambda s: min((s + 1) / t["warmup_steps"], 1.0))
    lw = t["loss_weights"]
    generator.train(); patch_d.train(); multibin_d.train()                       # train.py:504-506
    for step in (1, 2):
        real, lens = _tiny_batch(step, int(fx["B"]), int(fx["T"]), cfg.mel_channels)
        real, lens = real.to(DEV), lens.to(DEV)
        recon_pre, recon_post = generator(real, lens)                            # :524
        # ---- _train_discriminator :380-412 ----
        opt_d.zero_grad()
        rl, rm, _ = patch_d(real, lens, return_features=True)
        fl, fm = patch_d(recon_post.detach(), lens)
        loss_d = gan_loss.discriminator_loss(rl, fl, rm, fm)
        rl2, rm2, _ = multibin_d(real, lens, return_features=True)
        fl2, fm2 = multibin_d(recon_post.detach(), lens)
        loss_mbd = sum(gan_loss.discriminator_loss(r, f, rm2[0], fm2[0]) for r, f in zip(rl2, fl2)) / len(rl2)
        loss_d = loss_d + loss_mbd
        loss_d.backward()
        torch.nn.utils.clip_grad_norm_(d_params, 1.0)
        opt_d.step()
        # ---- _train_generator :414-501 ----
        opt_g.zero_grad()
        patch_d.eval(); multibin_d.eval()
        l_pre = recon_all(recon_pre, real, lens) + 0.25 * recon_group(recon_pre, real, lens)
        l_post = recon_all(recon_post, real, lens) + 0.25 * recon_group(recon_post, real, lens)
        gl, gm, gf = patch_d(recon_post, lens, return_features=True)
        gl2, gm2, gf2 = multibin_d(recon_post, lens, return_features=True)
        loss_gan = 0.5 * (gan_loss.generator_loss(gl, gm) + sum(gan_loss.generator_loss(g, gm2[0]) for g in gl2) / len(gl2))
        loss_fm = real.new_zeros(())
        if step == 2:                                                             # use_fm_loss :454-476
            with torch.no_grad():
                _, _, rf = patch_d(real, lens, return_features=True)
                _, _, rf2 = multibin_d(real, lens, return_features=True)
            fm_d1 = sum(masked_mae(ff, r, m) for (r, m), (ff, _) in zip(rf, gf)) / len(rf)
            fm_mbd = real.new_zeros(())
            for rfe, gfe in zip(rf2, gf2):
                for (r, m), (ff, _) in zip(rfe, gfe):
                    fm_mbd = fm_mbd + masked_mae(ff, r, m)
                fm_mbd = fm_mbd / len(rfe)
            loss_fm = 0.5 * (fm_d1 + fm_mbd / len(gf2))
        total = l_pre * 1.0 + l_post * 2.0 + loss_gan * lw["Gloss_lambda"] + loss_fm * lw["fm_lambda"]
        total.backward()
        torch.nn.utils.clip_grad_norm_(generator.parameters(), 1.0)
        opt_g.step()
        sched_g.step()
        got = np.array([float(loss_d), float(total), float(l_pre), float(l_post), float(loss_gan), float(loss_fm)])
        np.testing.assert_allclose(got, fx[f"s{step}_losses"], rtol=1e-3, atol=1e-6)
    ps = np.array([float(p.detach().double().sum()) for p in (dict(generator.named_parameters())[str(k)] for k in fx["g_keys"])])
    assert np.abs(ps - fx["s2_g_param_sums"]).max() < 2e-2
