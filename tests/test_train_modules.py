"""Drop-in training modules (mqgan_b200/discriminators.py, losses.py) against the training oracle, which is pinned
bit-identical to the reference's own classes (tests/test_train_oracle.py).  CPU tests: these modules are plain torch
outside the CUDA fast path."""
import numpy as np
import pytest
import torch

from mqgan_b200 import spec as S
from mqgan_b200.discriminators import MelSpectrogramPatchDiscriminator2D, MultiBinDiscriminator
from mqgan_b200.losses import LSGANLoss, MaskedMelLoss
from mqgan_b200.synth import synth_disc_state_dict, synth_lengths, synth_mels
from oracle import train_oracle as TO


def _batch(B=3, T=40, n_mels=32, seed=5):
    x = synth_mels(B, T, n_mels, seed=seed)
    lens = synth_lengths(B, T, seed=seed, ragged=True)
    return x.masked_fill((torch.arange(T)[None, :] >= lens[:, None]).unsqueeze(-1), 0.0), lens


def test_patch_discriminator_module_keys_and_arithmetic():
    dc = S.TINY_PATCH_D
    m = MelSpectrogramPatchDiscriminator2D(dc.mel_channels, list(dc.hidden_channels), [k[0] for k in dc.kernels],
                                           stride=[list(s) for s in dc.strides])
    assert sorted(m.state_dict()) == sorted(k for k, _ in S.patch_disc_param_spec(dc))
    assert {k for k, _ in m.named_buffers()} == {k for k, _ in S.patch_disc_param_spec(dc) if S.is_disc_buffer(k)}
    sd = synth_disc_state_dict(S.patch_disc_param_spec(dc), seed=7)
    m.load_state_dict(sd, strict=True)
    osd = {k: v.clone() for k, v in sd.items()}
    x, lens = _batch()
    ocfg = TO.patch_cfg([k[0] for k in dc.kernels], dc.strides)
    m.train()
    for _ in range(2):                                  # training mode: one power iteration per forward, buffers move
        out, mask, feats = m(x, lens, return_features=True)
        ro, rmask, rfeats = TO.patch_discriminator(osd, ocfg, x, lens, True)
        assert torch.allclose(out, ro, rtol=1e-5, atol=1e-6) and torch.equal(mask, rmask) and len(feats) == len(rfeats)
        assert torch.allclose(m.state_dict()["convs.1.weight_u"], osd["convs.1.weight_u"], atol=1e-7)
    m.eval()
    u_before = m.state_dict()["convs.1.weight_u"].clone()
    out, mask = m(x, lens)
    ro, rmask, _ = TO.patch_discriminator(osd, ocfg, x, lens, False)
    assert torch.allclose(out, ro, rtol=1e-5, atol=1e-6) and torch.equal(m.state_dict()["convs.1.weight_u"], u_before)
    # gradients reach weight_orig / bias / squeeze-excite, not the spectral-norm vectors
    m.train()
    out, mask = m(x, lens)
    (out * mask).pow(2).sum().backward()
    assert all(p.grad is not None for _, p in m.named_parameters())


def test_constructor_argument_forms():
    a = MelSpectrogramPatchDiscriminator2D(16, [8, 8], [5, 3, 3], stride=2)                    # int: time-only stride
    assert a.cfg.strides == ((1, 2), (1, 2), (1, 2)) and a.cfg.layer_stride(2) == (1, 1)
    b = MelSpectrogramPatchDiscriminator2D(16, [8, 8], [5, 3, 3], stride=(2, 2))
    assert b.cfg.strides == ((2, 2),) * 3
    c = MelSpectrogramPatchDiscriminator2D(16, [8, 8], [5, 3, 3], stride=(2, 2), lengthwise_only=True)
    assert c.cfg.kernels == ((1, 5), (1, 3), (1, 3)) and c.cfg.strides == ((1, 2),) * 3
    x, lens = _batch(2, 24, 16)
    out, mask = c(x, lens)
    assert out.shape[2] == 16 and out.shape[3] == 6                                             # mel axis untouched, time / 4
    with pytest.raises(AssertionError):
        MelSpectrogramPatchDiscriminator2D(16, [8, 8], [5, 3])
    with pytest.raises(AssertionError):
        MultiBinDiscriminator(30, 4)
    with pytest.raises(AssertionError):
        MultiBinDiscriminator(32, 4, hidden_channels=[6, 8], kernel_sizes=[3, 3, 3])


def test_multibin_module_matches_oracle():
    mc = S.TINY_MULTIBIN_D
    m = MultiBinDiscriminator(mc.mel_channels, mc.n_bins, list(mc.hidden_channels), list(mc.kernel_sizes), mc.n_no_strides)
    assert sorted(m.state_dict()) == sorted(k for k, _ in S.multibin_param_spec(mc))
    sd = synth_disc_state_dict(S.multibin_param_spec(mc), seed=8)
    m.load_state_dict(sd, strict=True)
    x, lens = _batch()
    m.eval()
    outs, masks, feats = m(x, lens, return_features=True)
    ro, rm, rf = TO.multibin_discriminator({k: v.clone() for k, v in sd.items()},
                                           TO.multibin_cfg(mc.kernel_sizes, mc.n_bins, mc.n_no_strides), x, lens, False)
    assert len(outs) == mc.n_bins == len(masks) == len(feats)
    for a, b, c, d in zip(outs, ro, masks, rm):
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()) and torch.equal(c, d)    # unsettled u, v: |logits| ~ 1e5
    assert len(m(x, lens)) == 2


def test_loss_modules_match_oracle():
    g = torch.Generator().manual_seed(0)
    real, fake = torch.randn(3, 1, 5, 7, generator=g), torch.randn(3, 1, 5, 7, generator=g)
    rmask = torch.rand(3, 1, 5, 7, generator=g) > 0.3
    fmask = torch.rand(3, 1, 5, 7, generator=g) > 0.3
    loss, st = LSGANLoss(), TO.LeCamState()
    for _ in range(3):                                   # EMA initialises on the first call, then decays
        a = loss.discriminator_loss(real, fake, rmask, fmask)
        b = TO.lsgan_d_loss(st, real, fake, rmask, fmask)
        assert torch.allclose(a, b, rtol=1e-6)
        assert torch.allclose(loss.ema_real, st.ema_real) and torch.allclose(loss.ema_fake, st.ema_fake)
        real, fake = real + 0.1, fake - 0.05
    assert torch.allclose(loss.generator_loss(fake, fmask), TO.masked_mse(fake, 1.0, fmask))
    assert float(loss.generator_loss(fake, torch.zeros_like(fmask))) == 0.0
    assert torch.allclose(LSGANLoss(use_lecam=False).discriminator_loss(real, fake), 0.5 * (((real - 1) ** 2).mean() + (fake ** 2).mean()))
    assert set(loss.state_dict()) == {"ema_real", "ema_fake"}
    x, lens = _batch()
    y = x + 0.3 * torch.randn(x.shape, generator=g)
    for gs in (1, 16):
        assert torch.allclose(MaskedMelLoss("mse", group_size=gs)(x, y, lens), TO.masked_mel_loss(x, y, lens, gs), rtol=1e-6)
    ch = MaskedMelLoss()(x, y, lens)                     # charbonnier, the class default
    pad = (torch.arange(x.shape[1])[None, :] >= lens[:, None])[:, :, None].expand_as(x)
    ref = torch.sqrt((x - y) ** 2 + 1e-12).masked_fill(pad, 0).sum(dim=(0, 1)) / (~pad).float().sum(dim=(0, 1))
    assert torch.allclose(ch, ref.mean(), rtol=1e-5)
    with pytest.raises(AssertionError):
        MaskedMelLoss("l1")


# ----------------------------------------------------------------------------
# space-to-depth lowering of strided convolutions (host math of mqgan_b200.training, checked against F.conv2d)
# ----------------------------------------------------------------------------
def _taps_conv_reference(x, weight, kind, taps):
    """Plain-torch stride-1 convolution over a channel-last image in the library's conventions."""
    import torch.nn.functional as F
    if kind == "linear":
        return x @ weight.t()
    if kind == "conv2d3":
        return F.conv2d(x.permute(0, 3, 1, 2), weight, padding=1).permute(0, 2, 3, 1)
    dh, dw = taps
    B, H, W, C = x.shape
    pad = max(max(abs(v) for v in dh), max(abs(v) for v in dw))
    xp = F.pad(x, (0, 0, pad, pad, pad, pad))
    y = 0
    for t, (a, b) in enumerate(zip(dh, dw)):
        y = y + xp[:, pad + a:pad + a + H, pad + b:pad + b + W, :] @ weight[:, :, t].t()
    return y


@pytest.mark.parametrize("kh,kw,sh,sw,H,W,cin", [
    (5, 5, 2, 2, 16, 12, 6), (5, 5, 1, 2, 9, 14, 1), (5, 5, 2, 2, 15, 11, 3),      # odd sizes: zero rows appended
    (3, 3, 2, 1, 8, 10, 4), (3, 3, 1, 2, 16, 9, 5), (3, 3, 1, 1, 7, 7, 2),
    (3, 7, 1, 1, 16, 20, 1), (3, 5, 1, 1, 16, 20, 8), (7, 7, 2, 2, 20, 16, 2), (7, 7, 1, 2, 10, 16, 1),
])
def test_strided_conv_lowering_equals_conv2d(kh, kw, sh, sw, H, W, cin):
    import torch.nn.functional as F
    from mqgan_b200 import training as TR
    g = torch.Generator().manual_seed(kh * 100 + sh * 10 + cin)
    cout = 5
    x = torch.randn(2, H, W, cin, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(cout, cin, kh, kw, generator=g, dtype=torch.float64, requires_grad=True)
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, stride=(sh, sw), padding=((kh - 1) // 2, (kw - 1) // 2)).permute(0, 2, 3, 1)
    got = TR.strided_conv_nhwc(x, w, (sh, sw), _taps_conv_reference)
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, atol=1e-10)
    dy = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    gx, gw = torch.autograd.grad(ref, (x, w), dy, retain_graph=True)
    hx, hw = torch.autograd.grad(got, (x, w), dy)
    assert torch.allclose(gx, hx, atol=1e-10) and torch.allclose(gw, hw, atol=1e-10)      # the lowering is differentiable
    if cin > 1 and (sh, sw) != (1, 1):
        dh, dw, index = TR.strided_conv_lowering(kh, kw, sh, sw)
        assert len(dh) <= 16 and len(set(index)) == kh * kw                               # fits MQ_MAX_TAPS; a bijection
