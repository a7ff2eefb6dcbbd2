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
