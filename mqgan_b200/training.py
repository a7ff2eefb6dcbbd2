"""Training step of the PreEncoder on B200s (BASELINE configs[4]).

Mirrors one iteration of the reference's loop - ``Trainer._train_epoch`` body (train.py:521-529):
generator forward (preencoder.py:363-418), ``_train_discriminator`` (:380-412), ``_train_generator``
(:414-501) - with the same losses (losses.py), discriminators (discriminators.py, legacy spectral norm),
Adam + warm-up (train.py:312-329) and gradient clipping, as data-parallel replicas: one process per GPU,
gradients averaged with NCCL all-reduces that are launched per bucket from autograd hooks while the
backward pass is still running (the reference has no multi-GPU training to be compatible with).

What runs where in this build (round 1 of the training variant):
  * every wide convolution / linear of the generator - forward, data gradient and weight gradient - runs
    in libmqgan_b200.so on tcgen05 (``mq_conv_gemm`` forward and, on the mirrored weight, data gradient;
    ``mq_conv_wgrad`` weight gradient): bf16 operands, fp32 accumulate, fp32 activations between layers.
    That is 99.7 % of the generator's FLOPs.
  * ConvBlock2D `pre` / `post` run in ``mq_cb2d_point_forward`` and ``mq_cb2d_backward`` so the
    (B, C, C, T) expansion of preencoder.py:288-295 never exists in memory, forward or backward.
  * the remaining element-wise / reduction work (APTx, masks, CBAM, pooling, FSQ straight-through, the
    1-channel stem / tail convolutions, the 4-wide quantiser projections), the discriminators (strided
    Conv2d: cuDNN) and Adam are PyTorch ops under autograd for now; DESIGN.md 3.7 lists them as the next
    kernels.
There is no CPU path: tensors must live on a CUDA device and the shared library must be present.

Dropout: the reference hard-wires p = 0.1 into parts of the generator regardless of its ``dropout``
argument (preencoder.py:109, 121, 233).  This step implements ``dropout = 0`` only (what the parity
fixtures pin); a non-zero ``dropout_p`` raises.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import ops
from .spec import MultiBinConfig, PatchDiscConfig, PreEncoderConfig, is_disc_buffer

Tensor = torch.Tensor


# ----------------------------------------------------------------------------
# convolutions on the tcgen05 kernels, differentiable
# ----------------------------------------------------------------------------
def _weight_from_taps(dw: Tensor, kind: str) -> Tensor:
    """(taps, cout, cin) -> the weight's own shape."""
    if kind == "linear":
        return dw[0]
    if kind in ("same1d", "causal1d", "taps2d"):
        return dw.permute(1, 2, 0)
    return dw.permute(1, 2, 0).reshape(dw.shape[1], dw.shape[2], 3, 3)


def _dgrad_launch(dyb: Tensor, weight: Tensor, kind: str, N: int, H: int, W: int, out: Tensor, tag: str, taps=None,
                  **epi) -> None:
    """dx = conv(dy, mirrored weight) into ``out`` (N, H, W, Cin), fp32 or bf16.  mq_conv_gemm takes at most 1024 output
    channels per launch, and a data gradient's output channels are the layer's INPUT channels (1152 for hifimusic's
    first up-block): wider ones are produced in channel slices."""
    if kind == "taps2d":
        wd, kd, dtaps = ops.dgrad_weight(weight, kind, taps)
    else:
        (wd, kd), dtaps = ops.dgrad_weight(weight, kind), None
    cin = wd.shape[0]
    key = "out_bf16" if out.dtype == torch.bfloat16 else "out_f32"
    off = "bf16_coff" if out.dtype == torch.bfloat16 else "f32_coff"
    step = cin if cin <= 1024 else 768
    for c0 in range(0, cin, step):
        pc = ops.pack_conv(wd[c0:c0 + step], None, kd, on_device=True, taps=dtaps)
        extra = {"res_coff": c0} if epi.get("res") is not None else {}
        ops.conv_gemm(dyb, pc, N, H, W, tag=tag, **{key: out, off: c0}, **epi, **extra)


class _ConvFn(torch.autograd.Function):
    """y = conv(x, w) + b on channel-last x (N, H, W, Cin); kind as ops.pack_conv."""

    @staticmethod
    def forward(ctx, x, weight, bias, kind: str, tag: str, taps=None, out_bf16: bool = False):
        N, H, W, cin = x.shape
        cout = weight.shape[0]
        xb = x.contiguous().to(torch.bfloat16)
        pc = ops.pack_conv(weight, bias, kind, on_device=True, taps=taps)
        out = torch.empty(N, H, W, cout, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=x.device)
        ops.conv_gemm(xb, pc, N, H, W, tag=tag, **({"out_bf16": out} if out_bf16 else {"out_f32": out}))
        ctx.save_for_backward(xb, weight)
        ctx.kind, ctx.tag, ctx.has_bias, ctx.x_dtype, ctx.taps = kind, tag, bias is not None, x.dtype, taps
        return out

    @staticmethod
    def backward(ctx, dy):
        xb, weight = ctx.saved_tensors
        N, H, W, cin = xb.shape
        cout = weight.shape[0]
        dyb = dy.contiguous().to(torch.bfloat16)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(N, H, W, cin, dtype=torch.bfloat16 if ctx.x_dtype == torch.bfloat16 else torch.float32,
                             device=dy.device)
            _dgrad_launch(dyb, weight, ctx.kind, N, H, W, dx, ctx.tag + ".dgrad", taps=ctx.taps)
            dx = dx.to(ctx.x_dtype)
        if ctx.needs_input_grad[1]:
            dh, dwt = ops.conv_taps(ctx.kind, weight.shape, ctx.taps)
            dw = _weight_from_taps(ops.conv_wgrad(dyb, xb, N, H, W, cout, cin, dh, dwt, tag=ctx.tag + ".wgrad"), ctx.kind)
            dw = dw.reshape(weight.shape)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(dim=(0, 1, 2), dtype=torch.float32)
        return dx, dw, db, None, None, None, None


def conv(x: Tensor, weight: Tensor, bias: Optional[Tensor], kind: str, tag: str = "", taps=None,
         out_bf16: bool = False) -> Tensor:
    """Channel-last convolution / linear through the tcgen05 kernels.  x: (N, H, W, Cin) fp32 or bf16 -> fp32 (bf16 with
    ``out_bf16``).  kind "taps2d": weight (Cout, Cin, ntaps) with ``taps`` = (dh list, dw list)."""
    if not x.is_cuda:
        raise RuntimeError("mqgan_b200.training runs on CUDA (B200) only - there is no CPU fallback")
    if kind == "linear" and weight.dim() == 3:
        weight = weight[:, :, 0]                                  # a 1x1 Conv1d (ResidualBlock1D.residual)
    cout = weight.shape[0]
    x, weight, bias = _pad_channels(x, weight, bias)              # TMA needs 16-byte channel pitches
    y = _ConvFn.apply(x, weight, bias, kind, tag, taps, out_bf16)
    return y[..., :cout] if weight.shape[0] != cout else y


def aptx(x: Tensor, beta, gamma) -> Tensor:
    """(1 + tanh(beta x)) * gamma * x (attentions.py:34-35)."""
    return (1 + torch.tanh(beta * x)) * gamma * x


# ----------------------------------------------------------------------------
# ConvBlock2D `pre` / `post` (preencoder.py:277-301), never expanding to (B, C, C, T)
# ----------------------------------------------------------------------------
class _Cb2dPointFn(torch.autograd.Function):
    """y[p] = sum_k wout_k * aptx(wpw_k * s[p] + bpw_k; 1, .5) + bout over valid rows, bout at padded rows."""

    @staticmethod
    def forward(ctx, s, wpw, bpw, wout, bout, row_mask, fast_tanh):
        y = ops.cb2d_point_forward(s, wpw, bpw, wout, bout, row_mask, fast_tanh)
        ctx.save_for_backward(s, wpw, bpw, wout, row_mask)
        ctx.fast_tanh = fast_tanh
        return y

    @staticmethod
    def backward(ctx, dy):
        s, wpw, bpw, wout, row_mask = ctx.saved_tensors
        ds, dwpw, dbpw, dwout, dbout = ops.cb2d_point_backward(s, dy.contiguous(), wpw, bpw, wout, row_mask, ctx.fast_tanh)
        return ds, dwpw, dbpw, dwout, dbout, None, None


def convblock2d(x: Tensor, mask_bt: Tensor, w: Dict[str, Tensor], prefix: str, native: bool = True,
                fast_tanh: bool = True) -> Tensor:
    """x (B, T, C) -> (B, T, C).  The 5x5 depth-wise conv over the (channel, time) plane is a 25-tap torch
    conv2d; the C-fold point-wise expansion + APTx + contraction is one fused kernel each way."""
    B, T, Cc = x.shape
    img = x.permute(0, 2, 1).unsqueeze(1)                                             # (B,1,C,T) view
    s = F.conv2d(img, w[prefix + ".dw.weight"], w[prefix + ".dw.bias"], padding=2)   # :286
    s = s.squeeze(1).permute(0, 2, 1).masked_fill(mask_bt.unsqueeze(-1), 0.0)        # (B,T,C) :287
    wpw = w[prefix + ".pw.weight"].reshape(Cc)
    bpw = w[prefix + ".pw.bias"].reshape(Cc)
    wout = w[prefix + ".conv_out.weight"].reshape(Cc)
    bout = w[prefix + ".conv_out.bias"].reshape(1)
    if native:
        return _Cb2dPointFn.apply(s.contiguous(), wpw, bpw, wout, bout, mask_bt.to(torch.uint8).contiguous(), fast_tanh)
    # plain-torch form of the same arithmetic (memory-hungry; kept as the on-device cross-check)
    u = (s.unsqueeze(-1) * wpw + bpw).masked_fill(mask_bt[:, :, None, None], 0.0)    # (B,T,C,K) :288-292
    return (aptx(u, 1.0, 0.5) * wout).sum(dim=-1) + bout                             # :293-295


# ----------------------------------------------------------------------------
# generator, training mode (channel-last)
# ----------------------------------------------------------------------------
def effective_weights(params: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """Fold both weight-norm flavours differentiably (w = g v / ||v||)."""
    out: Dict[str, Tensor] = {}
    for k, t in params.items():
        if k.endswith(".parametrizations.weight.original1"):
            base = k[: -len(".parametrizations.weight.original1")]
            out[base + ".weight"] = torch._weight_norm(t, params[base + ".parametrizations.weight.original0"], 0)
        elif k.endswith(".weight_v"):
            base = k[: -len(".weight_v")]
            out[base + ".weight"] = torch._weight_norm(t, params[base + ".weight_g"], 0)
        elif k.endswith("original0") or k.endswith(".weight_g"):
            continue
        else:
            out[k] = t
    return out


def _cbam(o: Tensor, mask_bt: Tensor, w: Dict[str, Tensor], prefix: str) -> Tensor:
    """CBAM1D with the reference's effective masking (attentions.py:248-273, 322-365, 411)."""
    mx = o.max(dim=1).values                                               # over ALL t
    valid = (~mask_bt).to(o.dtype).unsqueeze(-1)
    av = (o * valid).sum(dim=1) / valid.sum(dim=1).clamp(min=1.0)
    p = prefix + ".channel_attention.mlp."

    def mlp(v):
        return F.linear(F.relu(F.linear(v, w[p + "0.weight"], w[p + "0.bias"])), w[p + "2.weight"], w[p + "2.bias"])

    o1 = torch.sigmoid(mlp(mx) + mlp(av)).unsqueeze(1) * o
    pooled = torch.stack((o1.max(dim=2).values, o1.mean(dim=2)), dim=1)   # (B,2,T)
    logits = F.conv1d(pooled, w[prefix + ".spatial_attention.conv.weight"], None, padding=3)
    return torch.sigmoid(logits).permute(0, 2, 1) * o1 + o


def _residual_block(x: Tensor, mask_bt: Tensor, w: Dict[str, Tensor], prefix: str, causal: bool) -> Tensor:
    """ResidualBlock1D.forward (attentions.py:525-551), x (B, T, C)."""
    B, T, _ = x.shape
    beta, gamma = w[prefix + ".relu.beta"], w[prefix + ".relu.gamma"]
    kind = "causal1d" if causal else "same1d"
    m = mask_bt.unsqueeze(-1)
    x4 = x.reshape(B, T, 1, -1)
    if (prefix + ".residual.weight") in w:
        r = conv(x4, w[prefix + ".residual.weight"], w[prefix + ".residual.bias"], "linear", prefix + ".res").reshape(B, T, -1)
    else:
        r = x
    o = conv(x4, w[prefix + ".conv1.weight"], w[prefix + ".conv1.bias"], kind, prefix + ".conv1").reshape(B, T, -1)
    o = aptx(o.masked_fill(m, 0.0), beta, gamma)
    o = conv(o.reshape(B, T, 1, -1), w[prefix + ".conv2.weight"], w[prefix + ".conv2.bias"], kind, prefix + ".conv2").reshape(B, T, -1)
    if not causal:
        o = _cbam(o, mask_bt, w, prefix + ".cbam")
    return aptx((o + r).masked_fill(m, 0.0), beta, gamma)


def _pad_channels(x: Tensor, weight: Tensor, bias: Optional[Tensor]):
    """Zero-pad Cin / Cout to multiples of 8 (16-byte TMA channel pitches); differentiable."""
    cout, cin = weight.shape[0], weight.shape[1]
    pad_i, pad_o = (-cin) % 8, (-cout) % 8
    if pad_i:
        x = F.pad(x, (0, pad_i))
    if pad_i or pad_o:
        weight = F.pad(weight, (0, 0) * (weight.dim() - 2) + (0, pad_i, 0, pad_o))
    if pad_o and bias is not None:
        bias = F.pad(bias, (0, pad_o))
    return x, weight, bias


class _RefConvBlockFn(torch.autograd.Function):
    """ConvBlock.forward (preencoder.py:95-102) on a bf16 channel-last image whose padded rows are already zero:
    conv3x3 -> APTx -> conv3x3 -> APTx -> (+x) -> mask, as four library launches forward (two tcgen05 convs,
    two fused activation passes) and eight backward (two activation-gradient passes, two data-gradient and
    two weight-gradient convs, bias reductions in torch).  Saves the two fp32 pre-activations."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, mask8, residual: bool, tag: str):
        N, H, W, cin = x.shape
        c1 = w1.shape[0]
        u1 = torch.empty(N, H, W, c1, dtype=torch.float32, device=x.device)
        ops.conv_gemm(x, ops.pack_conv(w1, b1, "conv2d3", on_device=True), N, H, W, out_f32=u1, tag=tag + ".conv1")
        a1 = ops.act_forward(u1, None, None, W)
        u2 = torch.empty(N, H, W, w2.shape[0], dtype=torch.float32, device=x.device)
        ops.conv_gemm(a1, ops.pack_conv(w2, b2, "conv2d3", on_device=True), N, H, W, out_f32=u2, tag=tag + ".conv2")
        y = ops.act_forward(u2, x if residual else None, mask8, W)
        ctx.save_for_backward(x, u1, a1, u2, w1, w2, mask8)
        ctx.residual, ctx.tag = residual, tag
        return y

    @staticmethod
    def backward(ctx, dy):
        x, u1, a1, u2, w1, w2, mask8 = ctx.saved_tensors
        N, H, W, cin = x.shape
        c1, c2 = w1.shape[0], w2.shape[0]
        tag = ctx.tag
        dh, dwt = ops.conv_taps("conv2d3", w1.shape)
        du2, dres, db2 = ops.act_backward(dy.contiguous(), u2, mask8, W, want_res=ctx.residual, want_bias=True)
        dw2 = _weight_from_taps(ops.conv_wgrad(du2, a1, N, H, W, c2, c1, dh, dwt, tag=tag + ".conv2.wgrad"), "conv2d3")
        da1 = torch.empty(N, H, W, c1, dtype=torch.bfloat16, device=dy.device)
        _dgrad_launch(du2, w2, "conv2d3", N, H, W, da1, tag + ".conv2.dgrad")
        du1, _, db1 = ops.act_backward(da1, u1, None, W, want_bias=True)
        dw1 = _weight_from_taps(ops.conv_wgrad(du1, x, N, H, W, c1, cin, dh, dwt, tag=tag + ".conv1.wgrad"), "conv2d3")
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(N, H, W, cin, dtype=torch.bfloat16, device=dy.device)
            # the epilogue adds the skip path's gradient and re-applies the entry mask (x = x.masked_fill(mask) :96)
            _dgrad_launch(du1, w1, "conv2d3", N, H, W, dx, tag + ".conv1.dgrad", row_mask=mask8, mask_pre=True, res=dres,
                          res_mode=1 if dres is not None else 0)
        return dx, dw1, db1, dw2, db2, None, None, None


def _refiner_convblock(x: Tensor, mask_rows: Tensor, w: Dict[str, Tensor], prefix: str) -> Tensor:
    """x (B, T', F, C) bf16, zero at padded rows; mask_rows (B, T') bool."""
    w1, b1 = w[prefix + ".conv1.weight"], w[prefix + ".conv1.bias"]
    w2, b2 = w[prefix + ".conv2.weight"], w[prefix + ".conv2.bias"]
    residual = w1.shape[0] == w1.shape[1]
    x, w1, b1 = _pad_channels(x, w1, b1)
    return _RefConvBlockFn.apply(x.contiguous(), w1, b1, w2, b2, mask_rows.to(torch.uint8).contiguous(), residual, prefix)


def _refiner(r_in: Tensor, mask_bt: Tensor, w: Dict[str, Tensor], depth: int) -> Tensor:
    """UNetRefiner.forward (preencoder.py:169-202).  r_in (B, T, F) -> residual (B, T, mel).  Activations are
    bf16 between the blocks (fp32 accumulate and pre-activations inside them)."""
    B, T, Fw = r_in.shape
    mult = 1 << depth
    pad = (mult - T % mult) % mult
    cur = F.pad(mask_bt, (0, pad), value=True)                                      # pad_to_pow2_4d :29-47
    x = F.pad(r_in, (0, 0, 0, pad)).masked_fill(cur.unsqueeze(-1), 0.0).unsqueeze(-1).to(torch.bfloat16)   # (B,T8,F,1)
    x = _refiner_convblock(x, cur, w, "refiner.pre")
    skips: List[Tensor] = []
    for i in range(depth):
        skips.append(x)
        Bq, Tq, Fq, Cq = x.shape
        cur = cur.reshape(B, -1, 2).any(dim=2)                                      # max-pooled mask :65
        x = x.reshape(Bq, Tq // 2, 2, Fq, Cq).mean(dim=2).masked_fill(cur[:, :, None, None], 0.0)   # AvgPool2d((2,1)) :112, entry mask :96
        x = _refiner_convblock(x, cur, w, f"refiner.downs.{i}.conv")
    x = _refiner_convblock(x, cur, w, "refiner.mid")
    for i in range(depth):
        cur = cur.repeat_interleave(2, dim=1)                                       # nearest-upsampled mask :70
        # Upsample((2,1)) + cat :124-129; the up-sampled mask can cover a valid row of the skip tensor (odd lengths),
        # which ConvBlock's entry mask (:96) zeroes
        x = torch.cat([x.repeat_interleave(2, dim=1), skips.pop()], dim=-1).masked_fill(cur[:, :, None, None], 0.0)
        x = _refiner_convblock(x, cur, w, f"refiner.ups.{i}.conv")
    out = conv(x, w["refiner.post.weight"], w["refiner.post.bias"], "conv2d3", "refiner.post")   # x is masked already (:191)
    out = out.squeeze(-1)[:, :T, :].masked_fill(mask_bt.unsqueeze(-1), 0.0)         # :192-198
    return conv(out.reshape(B, T, 1, Fw), w["refiner.reproj.weight"], None, "linear", "refiner.reproj").reshape(B, T, -1)


_FSQ_CONST: Dict[tuple, tuple] = {}


def _fsq_constants(levels: Sequence[int], device) -> tuple:
    """(half_l, offset, shift, half_width) of quantizer.py:109-114, 132 as device tensors, built once per device (a
    host-to-device copy is illegal inside a CUDA-graph capture)."""
    key = (tuple(levels), str(device))
    if key not in _FSQ_CONST:
        lv = torch.tensor(list(levels), dtype=torch.int32, device=device)
        half_l = ((lv - 1) * (1 + 1e-3) / 2).float()
        offset = torch.where(lv % 2 == 0, 0.5, 0.0).float()
        _FSQ_CONST[key] = (half_l, offset, (offset / half_l).atanh(), (lv // 2).float())
    return _FSQ_CONST[key]


def fsq_quantize_ste(z: Tensor, levels: Sequence[int]) -> Tensor:
    """FSQ.quantize in training mode, noise_dropout = 0 (quantizer.py:109-114, 128-140): fp32, straight-through round."""
    half_l, offset, shift, half_w = _fsq_constants(levels, z.device)
    bounded = (z + shift).tanh() * half_l - offset
    return (bounded + (bounded.round() - bounded).detach()) / half_w


def generator_forward(params: Dict[str, Tensor], cfg: PreEncoderConfig, mel: Tensor, lengths: Tensor,
                      native_cb2d: bool = True, cb2d_fast_tanh: bool = True) -> Tuple[Tensor, Tensor]:
    """PreEncoder.forward (preencoder.py:363-418), dropout 0 -> (x_recon, x_post), both (B, T, mel)."""
    w = effective_weights(params)
    B, T, n_mels = mel.shape
    mask_bt = torch.arange(T, device=mel.device)[None, :] >= lengths.to(mel.device)[:, None]
    x = conv(mel.reshape(B, T, 1, n_mels), w["proj.weight"], w["proj.bias"], "linear", "proj").reshape(B, T, -1)
    x = convblock2d(x, mask_bt, w, "pre", native_cb2d, cb2d_fast_tanh)
    for i in range(len(cfg.encoder_layers)):
        x = _residual_block(x, mask_bt, w, f"encoder_blocks.{i}", causal=False)
    z = F.linear(x, w["q_in_proj.weight"], w["q_in_proj.bias"])
    codes = fsq_quantize_ste(z.float(), cfg.fsq_levels)
    dec = F.linear(codes, w["q_out_proj.weight"], w["q_out_proj.bias"])
    for i in range(len(cfg.decoder_layers)):
        dec = _residual_block(dec, mask_bt, w, f"decoder_blocks.{i}", causal=True)
    xr = convblock2d(dec, mask_bt, w, "post", native_cb2d, cb2d_fast_tanh)
    x_recon = conv(xr.reshape(B, T, 1, -1), w["out_proj.weight"], w["out_proj.bias"], "linear", "out_proj").reshape(B, T, -1)
    hid = conv(dec.reshape(B, T, 1, -1), w["hidden_proj.weight"], w["hidden_proj.bias"], "linear", "hidden_proj").reshape(B, T, -1)
    r_in = torch.cat([x_recon, hid], dim=2).detach()                               # :411-413
    return x_recon, x_recon + _refiner(r_in, mask_bt, w, cfg.refiner_depth)


# ----------------------------------------------------------------------------
# strided convolutions on the stride-1 tcgen05 kernels: space-to-depth lowering (host math, any device)
# ----------------------------------------------------------------------------
_LOWERING: Dict[tuple, tuple] = {}
_LOWERING_INDEX: Dict[tuple, Tensor] = {}


def strided_conv_lowering(kh: int, kw: int, sh: int, sw: int):
    """A (kh, kw) convolution with stride (sh, sw) and padding ((kh-1)//2, (kw-1)//2) equals a STRIDE-1 convolution over
    the space-to-depth image x'[h', w', (rh, rw, c)] = x[h' sh + rh, w' sw + rw, c]: input row h' sh + di (di = i - ph)
    is row h' + floor(di / sh) of x', phase di mod sh.  Returns (dh, dw, index): the row-major tap offsets of the
    lowered kernel and, for every original tap (i, j) in row-major order, its slot tap' * (sh sw) + rh sw + rw in the
    lowered weight (slots that receive no tap stay zero: a 5x5 stride-2 kernel becomes 3x3 over 4C channels, 25 of
    36 slots used)."""
    key = (kh, kw, sh, sw)
    if key not in _LOWERING:
        ph, pw = (kh - 1) // 2, (kw - 1) // 2
        rows = sorted({(i - ph) // sh for i in range(kh)})
        cols = sorted({(j - pw) // sw for j in range(kw)})
        dh = [a for a in rows for _ in cols]
        dw = [b for _ in rows for b in cols]
        index = []
        for i in range(kh):
            for j in range(kw):
                a, rh = divmod(i - ph, sh)
                b, rw = divmod(j - pw, sw)
                index.append((rows.index(a) * len(cols) + cols.index(b)) * (sh * sw) + rh * sw + rw)
        _LOWERING[key] = (dh, dw, index)
    return _LOWERING[key]


def space_to_depth(x: Tensor, sh: int, sw: int) -> Tensor:
    """(B, H, W, C) -> (B, ceil(H/sh), ceil(W/sw), sh sw C), channel (rh sw + rw) C + c; zero rows / columns are appended
    when H, W are not multiples of the stride."""
    if sh == 1 and sw == 1:
        return x
    B, H, W, C = x.shape
    eh, ew = (-H) % sh, (-W) % sw
    if eh or ew:
        x = F.pad(x, (0, 0, 0, ew, 0, eh))
    return (x.reshape(B, (H + eh) // sh, sh, (W + ew) // sw, sw, C).permute(0, 1, 3, 2, 4, 5)
            .reshape(B, (H + eh) // sh, (W + ew) // sw, sh * sw * C))


def lower_strided_weight(w: Tensor, sh: int, sw: int):
    """(Cout, Cin, kh, kw) -> ((Cout, sh sw Cin, ntaps') weight of the stride-1 convolution over space_to_depth(x),
    (dh, dw)); differentiable (a scatter of the original taps into a zero tensor)."""
    cout, cin, kh, kw = w.shape
    dh, dw, index = strided_conv_lowering(kh, kw, sh, sw)
    src = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    slots = len(dh) * sh * sw
    ikey = (kh, kw, sh, sw, str(w.device))
    if ikey not in _LOWERING_INDEX:                        # built once per device (no host-to-device copy inside a graph capture)
        _LOWERING_INDEX[ikey] = torch.tensor(index, dtype=torch.long, device=w.device)
    dst = w.new_zeros(cout, slots, cin).index_copy(1, _LOWERING_INDEX[ikey], src)
    return dst.reshape(cout, len(dh), sh * sw * cin).permute(0, 2, 1), (dh, dw)


def im2col_1ch(x: Tensor, kh: int, kw: int, sh: int, sw: int) -> Tensor:
    """(B, H, W, 1) -> (B, Ho, Wo, kh kw) patches of a single-channel image (the discriminators' first layers:
    with one input channel the convolution is a linear map of the kh kw patch)."""
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    xp = F.pad(x[..., 0], (pw, pw, ph, ph))
    return xp.unfold(1, kh, sh).unfold(2, kw, sw).reshape(x.shape[0], -1, (x.shape[2] + 2 * pw - kw) // sw + 1, kh * kw)


def strided_conv_nhwc(x: Tensor, w: Tensor, stride: Tuple[int, int], conv_fn) -> Tensor:
    """conv2d(x, w, stride, padding=((kh-1)//2, (kw-1)//2)) on a channel-last image through a stride-1 convolution
    ``conv_fn(x', weight, kind, taps)`` (the tcgen05 kernels on the GPU; tests pass a plain-torch one)."""
    cout, cin, kh, kw = w.shape
    sh, sw = stride
    if cin == 1:
        return conv_fn(im2col_1ch(x, kh, kw, sh, sw), w.reshape(cout, kh * kw), "linear", None)
    if (sh, sw) == (1, 1) and (kh, kw) == (3, 3):
        return conv_fn(x, w, "conv2d3", None)
    wl, (dh, dw) = lower_strided_weight(w, sh, sw)
    xs = space_to_depth(x, sh, sw)
    if dh == [-1, -1, -1, 0, 0, 0, 1, 1, 1] and dw == [-1, 0, 1] * 3:
        return conv_fn(xs, wl.reshape(cout, wl.shape[1], 3, 3), "conv2d3", None)
    return conv_fn(xs, wl.contiguous(), "taps2d", (dh, dw))


# ----------------------------------------------------------------------------
# discriminators (discriminators.py), legacy spectral norm
# ----------------------------------------------------------------------------
def _spectral_weight(sd: Dict[str, Tensor], prefix: str, training: bool) -> Tensor:
    """torch.nn.utils.spectral_norm: one in-place power iteration of (u, v) per training-mode forward, none
    in eval; w = w_orig / (u . W v)."""
    w_orig = sd[prefix + ".weight_orig"]
    u, v = sd[prefix + ".weight_u"], sd[prefix + ".weight_v"]
    wm = w_orig.reshape(w_orig.shape[0], -1)
    if training:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=1e-12))
            u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=1e-12))
        u, v = u.clone(), v.clone()
    return w_orig / torch.dot(u, torch.mv(wm, v))


class _LeakyMaskFn(torch.autograd.Function):
    """LeakyReLU(0.2)(y + bias) with the padded patches zeroed (discriminators.py:234, 247) in one pass each way, bf16 out.
    The convolution in front runs WITHOUT its bias: the add happens here, and the bias gradient is the column sum the
    backward pass produces on the fly (cuDNN would re-read the whole gradient tensor for it)."""

    @staticmethod
    def forward(ctx, y, bias, pix_mask):
        ctx.save_for_backward(y, bias, pix_mask)
        return ops.leaky_mask_forward(y, pix_mask, 0.2, bias)

    @staticmethod
    def backward(ctx, dout):
        y, bias, pix_mask = ctx.saved_tensors
        dout = dout.contiguous(memory_format=torch.channels_last)
        if ctx.needs_input_grad[1]:
            du, db = ops.leaky_mask_backward(dout, y, pix_mask, 0.2, bias, want_bias=True)
        else:                                                  # generator step: the discriminator is only a loss
            du, db = ops.leaky_mask_backward(dout, y, pix_mask, 0.2, bias), None
        return du.to(y.dtype), db, None


def patch_discriminator(sd: Dict[str, Tensor], dc: PatchDiscConfig, x: Tensor, lengths: Tensor, training: bool,
                        prefix: str = "", autocast_bf16: bool = False, native_conv: bool = False):
    """MelSpectrogramPatchDiscriminator2D.forward (discriminators.py:208-257): x (B, T, F) ->
    (logits (B,1,H,W), valid-patch mask, [(feature, padded mask)]).

    ``autocast_bf16`` is the reference's CUDA training precision (train.py:523: convs under bf16 autocast, the
    activations after them bf16): feature maps stay bf16 and channels_last from layer to layer, and the
    LeakyReLU + patch mask after each conv is one library pass.  Otherwise everything is fp32 (parity mode).

    ``native_conv`` (with ``autocast_bf16``): the convolutions themselves run on the library's stride-1 tcgen05 kernels
    instead of cuDNN - strided layers through the space-to-depth lowering (``strided_conv_nhwc``), one-channel first
    layers as im2col + linear - with forward, data and weight gradients as for the generator's convolutions."""
    if native_conv and autocast_bf16 and x.is_cuda:
        return _patch_discriminator_native(sd, dc, x, lengths, training, prefix)
    B, T, Fm = x.shape
    n = len(dc.kernels)
    pad_mask = (torch.arange(T, device=x.device)[None, :] >= lengths.to(x.device)[:, None])[:, None, None, :].expand(-1, 1, Fm, -1)
    out = x.transpose(1, 2).unsqueeze(1)
    if autocast_bf16:
        out = out.to(torch.bfloat16)
    feats = []
    for i in range(n):
        if i == n - 1:                                                              # masked squeeze-excite :10-67, :231-232
            valid = ~pad_mask
            denom = valid.sum(dim=(2, 3)).clamp(min=1)
            sq = (out.float() * valid).sum(dim=(2, 3)) / denom
            h = F.relu(F.linear(sq, sd[prefix + "se_block.fc1.weight"], sd[prefix + "se_block.fc1.bias"]))
            ex = torch.sigmoid(F.linear(h, sd[prefix + "se_block.fc2.weight"], sd[prefix + "se_block.fc2.bias"]))
            out = (out * ex.reshape(B, -1, 1, 1)).to(out.dtype)
        kh, kw = dc.kernels[i]
        sh, sw = dc.layer_stride(i)
        wgt = _spectral_weight(sd, f"{prefix}convs.{i}", training)
        bias = sd[f"{prefix}convs.{i}.bias"]
        if autocast_bf16:
            wgt, bias = wgt.to(torch.bfloat16), bias.to(torch.bfloat16)
        # channels_last on purpose: in NCHW / bf16 cuDNN's heuristic sends the data gradient of the (1, 2)-strided
        # 256 -> 384 multi-bin layer to dgrad2d_grouped_direct_kernel - 9.1 ms per call instead of 0.27 ms, 2/3 of the
        # whole step (tools/disc_conv_probe.py, profiles/disc_conv_probe_r01.log)
        fused = autocast_bf16 and wgt.shape[0] % 8 == 0                             # library activation pass (adds the bias itself)
        y = F.conv2d(out.contiguous(memory_format=torch.channels_last), wgt.contiguous(memory_format=torch.channels_last),
                     None if fused else bias, stride=(sh, sw), padding=((kh - 1) // 2, (kw - 1) // 2))
        if sh > 1 or sw > 1:
            pad_mask = F.max_pool2d(pad_mask.float(), kernel_size=(sh, sw), stride=(sh, sw), ceil_mode=True).bool()
        if fused:
            out = _LeakyMaskFn.apply(y.contiguous(memory_format=torch.channels_last), sd[f"{prefix}convs.{i}.bias"],
                                     pad_mask.reshape(B, y.shape[2], y.shape[3]).to(torch.uint8).contiguous())
        else:
            out = F.leaky_relu(y, 0.2).masked_fill(pad_mask, 0.0)
        if dc.feature_layers[i]:
            feats.append((out, pad_mask))
    return out.float(), ~pad_mask, feats


def _patch_discriminator_native(sd: Dict[str, Tensor], dc: PatchDiscConfig, x: Tensor, lengths: Tensor, training: bool,
                                prefix: str):
    """patch_discriminator with channel-last bf16 feature maps (B, H, W, C) and every convolution on the tcgen05 kernels."""
    B, T, Fm = x.shape
    n = len(dc.kernels)
    pad_mask = (torch.arange(T, device=x.device)[None, :] >= lengths.to(x.device)[:, None])[:, None, None, :].expand(-1, 1, Fm, -1)
    out = x.transpose(1, 2).unsqueeze(-1).to(torch.bfloat16)                        # (B, F, T, 1)
    feats = []
    for i in range(n):
        if i == n - 1:                                                              # masked squeeze-excite
            valid = (~pad_mask).permute(0, 2, 3, 1)                                 # (B, H, W, 1)
            denom = valid.sum(dim=(1, 2)).clamp(min=1)
            sq = (out.float() * valid).sum(dim=(1, 2)) / denom
            h = F.relu(F.linear(sq, sd[prefix + "se_block.fc1.weight"], sd[prefix + "se_block.fc1.bias"]))
            ex = torch.sigmoid(F.linear(h, sd[prefix + "se_block.fc2.weight"], sd[prefix + "se_block.fc2.bias"]))
            out = (out * ex[:, None, None, :]).to(torch.bfloat16)
        sh, sw = dc.layer_stride(i)
        wgt = _spectral_weight(sd, f"{prefix}convs.{i}", training)
        bias = sd[f"{prefix}convs.{i}.bias"]
        tag = f"{prefix}convs.{i}"
        y = strided_conv_nhwc(out, wgt, (sh, sw),
                              lambda xx, ww, kind, taps: conv(xx, ww, None, kind, tag, taps=taps, out_bf16=True))
        if sh > 1 or sw > 1:
            pad_mask = F.max_pool2d(pad_mask.float(), kernel_size=(sh, sw), stride=(sh, sw), ceil_mode=True).bool()
        if wgt.shape[0] % 8 == 0:
            out = _LeakyMaskFn.apply(y.permute(0, 3, 1, 2), bias,
                                     pad_mask.reshape(B, y.shape[1], y.shape[2]).to(torch.uint8).contiguous()).permute(0, 2, 3, 1)
        else:                                                                       # the 1-channel logits layer
            out = F.leaky_relu(y.float() + bias, 0.2).masked_fill(pad_mask.permute(0, 2, 3, 1), 0.0)
        if dc.feature_layers[i]:
            feats.append((out.permute(0, 3, 1, 2), pad_mask))
    return out.permute(0, 3, 1, 2).float(), ~pad_mask, feats


_SIDE_STREAMS: Dict[str, List["torch.cuda.Stream"]] = {}


def _side_streams(device, n: int) -> List["torch.cuda.Stream"]:
    key = str(device)
    pool = _SIDE_STREAMS.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


def multibin_discriminator(sd: Dict[str, Tensor], mc: MultiBinConfig, x: Tensor, lengths: Tensor, training: bool,
                           autocast_bf16: bool = False, concurrent: bool = True, native_conv: bool = False):
    """MultiBinDiscriminator.forward (discriminators.py:292-312).  The bands are independent networks on 1/n_bins of
    the mel axis, each too small to fill 148 SMs: with ``concurrent`` every band runs on its own CUDA stream (forked
    from / joined to the caller's stream; autograd replays each band's backward on the same stream, and a CUDA-graph
    capture records them as parallel branches)."""
    subs = torch.split(x, x.size(-1) // mc.n_bins, dim=-1)
    res = []
    if concurrent and x.is_cuda and mc.n_bins > 1:
        cur = torch.cuda.current_stream(x.device)
        streams = _side_streams(x.device, mc.n_bins)
        for b, sub in enumerate(subs):
            streams[b].wait_stream(cur)
            with torch.cuda.stream(streams[b]):
                res.append(patch_discriminator(sd, mc.bin_config, sub, lengths, training, f"discriminators.{b}.", autocast_bf16,
                                               native_conv))
        for st in streams:
            cur.wait_stream(st)
    else:
        for b, sub in enumerate(subs):
            res.append(patch_discriminator(sd, mc.bin_config, sub, lengths, training, f"discriminators.{b}.", autocast_bf16,
                                           native_conv))
    return [r[0] for r in res], [r[1] for r in res], [r[2] for r in res]


# ----------------------------------------------------------------------------
# losses (losses.py, train.py:38-45)
# ----------------------------------------------------------------------------
def masked_mse(pred: Tensor, target: float, mask: Tensor) -> Tensor:
    """LSGANLoss._masked_mse (losses.py:21-35), mask True = valid; no host sync: an empty mask gives 0."""
    m = mask.to(pred.dtype)
    valid = m.sum()
    return torch.where(valid > 0, (((pred - target) ** 2) * m).sum() / valid.clamp(min=1), pred.new_zeros(()))


def masked_mel_loss(x: Tensor, y: Tensor, lengths: Tensor, group_size: int) -> Tensor:
    """MaskedMelLoss("mse", group_size) (losses.py:148-182)."""
    B, T, Cc = x.shape
    G = Cc // group_size
    pad = (torch.arange(T, device=x.device)[None, :] >= lengths.to(x.device)[:, None])[:, :, None].expand(B, T, Cc)
    pad = pad.reshape(B, T, G, group_size)
    per = ((x - y) ** 2).reshape(B, T, G, group_size).masked_fill(pad, 0.0)
    return (per.sum(dim=[0, 1, 3]) / ((~pad).to(x.dtype).sum(dim=[0, 1, 3]) + 1e-12)).mean()


def masked_mae(pred: Tensor, target: Tensor, mask: Tensor, eps: float = 1e-8) -> Tensor:
    """train.py:38-45, mask True = padded."""
    mask = mask.expand_as(pred)
    return (pred - target).abs().masked_fill(mask, 0.0).sum() / ((~mask).sum() + eps)


class LeCam:
    """EMA anchors of LSGANLoss (losses.py:17-20, 37-79).  Under data parallelism the batch means are
    averaged over the replicas before they enter the EMA, so every replica carries the same anchors."""

    def __init__(self, device, decay: float = 0.99, group=None):
        self.decay, self.group = decay, group
        self.ema = torch.zeros(2, device=device)          # [real, fake]
        self.initialized = False

    def d_loss(self, real: Tensor, fake: Tensor, real_mask: Tensor, fake_mask: Tensor) -> Tensor:
        """LSGANLoss.discriminator_loss (losses.py:81-107): the EMA moves first, then LeCam reads it."""
        loss = 0.5 * (masked_mse(real, 1.0, real_mask) + masked_mse(fake, 0.0, fake_mask))
        rm, fm = real_mask.to(real.dtype), fake_mask.to(fake.dtype)
        means = torch.stack(((real * rm).sum() / rm.sum().clamp(min=1), (fake * fm).sum() / fm.sum().clamp(min=1))).detach()
        if self.group is not None or (torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1):
            torch.distributed.all_reduce(means, group=self.group)
            means = means / torch.distributed.get_world_size(self.group)
        if not self.initialized:                          # in place: the buffer's address is baked into a captured CUDA graph
            self.ema.copy_(means)
            self.initialized = True
        else:
            self.ema.mul_(self.decay).add_(means, alpha=1 - self.decay)
        term_r = (((real - self.ema[1]).clamp(min=0) * rm) ** 2).sum() / rm.sum().clamp(min=1)
        term_f = (((self.ema[0] - fake).clamp(min=0) * fm) ** 2).sum() / fm.sum().clamp(min=1)
        return loss + term_r + term_f


# ----------------------------------------------------------------------------
# gradient all-reduce, bucketed and overlapped with the backward pass
# ----------------------------------------------------------------------------
class GradBucketReducer:
    """Flat gradient buckets whose all-reduce (average) starts as soon as the last gradient of the bucket
    has been accumulated, while autograd is still producing the others (NCCL runs on its own stream).

    Parameters keep ``.grad`` as VIEWS into the bucket, so nothing is copied either way; ``zero()`` clears
    the buckets instead of ``optimizer.zero_grad()``.  Buckets are filled in reverse registration order
    (gradients arrive roughly last-layer-first).  With world size 1 (or no process group) it only provides
    the flat buffers.  Parameters that receive no gradient in a step (hidden_proj, preencoder.py:411-413)
    would stall their bucket, so ``finish()`` launches whatever has not been launched."""

    def __init__(self, params: Sequence[Tensor], bucket_bytes: int = 25 << 20, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        # CUDA streams other than the hook's current one on which gradients of a bucket may have been produced (the
        # multi-bin discriminator's bands run on side streams): the all-reduce must wait for all of them, not only for
        # the stream the LAST gradient of the bucket arrived on
        self.producer_streams: List["torch.cuda.Stream"] = []
        self.world = torch.distributed.get_world_size(group) if torch.distributed.is_initialized() else 1
        self.buckets: List[Tensor] = []
        self._bucket_of: Dict[int, int] = {}
        self._pending: List[int] = []
        self._count: List[int] = []
        self._launched: List[bool] = []
        self._handles = []
        cur: List[Tensor] = []
        size = 0
        groups: List[List[Tensor]] = []
        for p in reversed(self.params):
            if cur and (size + p.numel()) * 4 > bucket_bytes:
                groups.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += p.numel()
        if cur:
            groups.append(cur)
        for bi, g in enumerate(groups):
            flat = torch.zeros(sum(p.numel() for p in g), dtype=torch.float32, device=g[0].device)
            off = 0
            for p in g:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
                self._bucket_of[id(p)] = bi
                if self.world > 1:
                    p.register_post_accumulate_grad_hook(self._hook)
            self.buckets.append(flat)
            self._count.append(len(g))
        self.zero()

    def zero(self):
        for b in self.buckets:
            b.zero_()
        self._pending = list(self._count)
        self._launched = [False] * len(self.buckets)
        self._handles = []

    def _launch(self, bi: int):
        self._launched[bi] = True
        if self.producer_streams and self.buckets[bi].is_cuda:
            cur = torch.cuda.current_stream(self.buckets[bi].device)
            for st in self.producer_streams:
                if st != cur:
                    cur.wait_stream(st)
        self._handles.append(torch.distributed.all_reduce(self.buckets[bi], group=self.group, async_op=True))

    def _hook(self, p: Tensor):
        bi = self._bucket_of[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0 and not self._launched[bi]:
            self._launch(bi)

    def finish(self):
        """Wait for every bucket's all-reduce and turn the sums into averages."""
        if self.world <= 1:
            return
        for bi in range(len(self.buckets)):
            if not self._launched[bi]:
                self._launch(bi)
        for h in self._handles:
            h.wait()
        for b in self.buckets:
            b.div_(self.world)


# ----------------------------------------------------------------------------
# the step
# ----------------------------------------------------------------------------
class TrainStep:
    """One replica of the reference's training iteration.

    ``g_sd`` / ``pd_sd`` / ``mb_sd``: state-dicts with the reference's key names (generator; patch and
    multi-bin discriminators incl. the spectral-norm ``weight_u`` / ``weight_v`` buffers).  ``tcfg``: the
    ``training`` section of a reference model_config*.yaml.  Call ``step(real, lengths)`` per batch; it returns
    the dict of losses the reference logs (train.py:494-500, plus ``loss_d``)."""

    def __init__(self, cfg: PreEncoderConfig, pd_cfg: PatchDiscConfig, mb_cfg: MultiBinConfig, g_sd, pd_sd, mb_sd,
                 tcfg: dict, device, dropout_p: float = 0.0, d_autocast_bf16: bool = False, native_cb2d: bool = True,
                 cb2d_fast_tanh: bool = True, d_native: bool = False, group=None):
        if dropout_p != 0.0:
            raise NotImplementedError("mqgan_b200.training implements dropout = 0 only (see module docstring)")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("mqgan_b200.training runs on CUDA (B200) only - there is no CPU fallback")
        ops._lib.lib()                                    # fail now if the extension is missing
        self.cfg, self.pd_cfg, self.mb_cfg, self.tcfg = cfg, pd_cfg, mb_cfg, tcfg
        self.device, self.group = dev, group
        self.d_autocast_bf16, self.native_cb2d, self.cb2d_fast_tanh = d_autocast_bf16, native_cb2d, cb2d_fast_tanh
        self.d_native = bool(d_native) and d_autocast_bf16          # discriminator convolutions on the tcgen05 kernels too
        self.g = {k: v.detach().clone().float().to(dev).requires_grad_(True) for k, v in g_sd.items()}
        self.pd = {k: v.detach().clone().float().to(dev) for k, v in pd_sd.items()}
        self.mb = {k: v.detach().clone().float().to(dev) for k, v in mb_sd.items()}
        for sd in (self.pd, self.mb):
            for k, v in sd.items():
                if not is_disc_buffer(k):
                    v.requires_grad_(True)
        self.d_training = True
        self.lecam = LeCam(dev, group=group)
        t = tcfg
        # learning rates live in device tensors and Adam is capturable, so the whole step can be replayed from a CUDA
        # graph (capture()); the warm-up schedule (train.py:326-329, LambdaLR) is applied to the tensor before each step
        self.lr_g = torch.tensor(float(t["lr"]), device=dev)
        self.lr_d = torch.tensor(float(t["lr"] * t["lr_d_factor"]), device=dev)
        self.opt_g = torch.optim.Adam(list(self.g.values()), lr=self.lr_g, betas=(t["beta1"], t["beta2"]), fused=True,
                                      capturable=True)
        self.opt_d = torch.optim.Adam(self.d_params(), lr=self.lr_d, betas=(t["d_beta1"], t["d_beta2"]), fused=True,
                                      capturable=True)
        self.g_steps = 0
        self._graphs: Dict[tuple, tuple] = {}
        _fsq_constants(cfg.fsq_levels, dev)
        self.red_g = GradBucketReducer(list(self.g.values()), group=group)
        self.red_d = GradBucketReducer(self.d_params(), group=group)
        self._band_streams = _side_streams(dev, mb_cfg.n_bins)

    def d_params(self) -> List[Tensor]:
        return [v for sd in (self.pd, self.mb) for v in sd.values() if v.requires_grad]

    def _patch(self, x, lengths, training):
        return patch_discriminator(self.pd, self.pd_cfg, x, lengths, training, autocast_bf16=self.d_autocast_bf16,
                                   native_conv=self.d_native)

    def _multibin(self, x, lengths, training):
        return multibin_discriminator(self.mb, self.mb_cfg, x, lengths, training, self.d_autocast_bf16,
                                      native_conv=self.d_native)

    def start_epoch(self):
        """train.py:504-506: the discriminators go back to training mode (power iteration on) each epoch."""
        self.d_training = True

    def generator_state_dict(self) -> Dict[str, Tensor]:
        return {k: v.detach() for k, v in self.g.items()}

    def rebind_lr(self) -> None:
        """Point both optimisers' ``lr`` back at the live device tensors.  ``Optimizer.load_state_dict`` replaces
        ``param_groups[*]['lr']`` with a copy, after which ``lr_g.fill_()`` would no longer reach Adam (and a later
        graph capture would bake the stale value in)."""
        for grp in self.opt_g.param_groups:
            grp["lr"] = self.lr_g
        for grp in self.opt_d.param_groups:
            grp["lr"] = self.lr_d

    def current_lr_g(self) -> float:
        """LambdaLR(min((step + 1) / warmup_steps, 1)) of train.py:326-329 for the step about to run."""
        t = self.tcfg
        return float(t["lr"]) * min((self.g_steps + 1) / t["warmup_steps"], 1.0)

    @ops.on_device
    def step(self, real: Tensor, lengths: Tensor, gan: bool = True, use_fm: Optional[bool] = None) -> Dict[str, Tensor]:
        """train.py:521-529 for one batch.  ``gan``: epoch >= discriminator_train_start_epoch.  Losses are
        returned as 0-d device tensors (one host read at the caller's discretion)."""
        self.lr_g.fill_(self.current_lr_g())
        out = self._step_body(real, lengths, gan, use_fm)
        self.g_steps += 1
        return out

    @ops.on_device
    def capture(self, real: Tensor, lengths: Tensor, gan: bool = True, use_fm: Optional[bool] = None, warmup: int = 3):
        """Record the whole iteration for this batch shape into a CUDA graph (≈ 7 000 launches become one replay).
        Runs ``warmup`` real iterations on a side stream first (they train: LeCam / spectral-norm state must be past
        their first-call branches), then captures.  Use ``step_graphed`` afterwards."""
        key = (tuple(real.shape), bool(gan), use_fm)
        real = real.to(self.device)
        lengths = lengths.to(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.last_losses = {k: v.clone() for k, v in self.step(real, lengths, gan, use_fm).items()}
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        s_real, s_len = real.clone(), lengths.clone()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            outs = self._step_body(s_real, s_len, gan, use_fm)
        self._graphs[key] = (graph, s_real, s_len, outs, self.last_recon)
        return key

    @ops.on_device
    def step_graphed(self, real: Tensor, lengths: Tensor, gan: bool = True, use_fm: Optional[bool] = None) -> Dict[str, Tensor]:
        """Replay the captured iteration on a new batch of the captured shape.  The returned loss tensors are the
        graph's static outputs (overwritten by the next replay)."""
        key = (tuple(real.shape), bool(gan), use_fm)
        if key not in self._graphs:
            raise KeyError(f"no CUDA graph captured for batch shape {tuple(real.shape)}; call capture() first")
        graph, s_real, s_len, outs, recon = self._graphs[key]
        s_real.copy_(real, non_blocking=True)
        s_len.copy_(lengths, non_blocking=True)
        self.lr_g.fill_(self.current_lr_g())
        graph.replay()
        self.g_steps += 1
        self.last_recon = recon
        return outs

    def _mark(self, name: str) -> None:
        """Phase boundary for tools/train_phases.py (CUDA event on the step's stream); a no-op unless a list is attached."""
        ev = getattr(self, "phase_events", None)
        if ev is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.append((name, e))

    def _step_body(self, real: Tensor, lengths: Tensor, gan: bool, use_fm: Optional[bool]) -> Dict[str, Tensor]:
        t = self.tcfg
        clip = t.get("clip_grad_norm", 1.0)
        lw = t["loss_weights"]
        use_fm = t.get("use_fm_loss", False) if use_fm is None else use_fm
        real = real.to(self.device, non_blocking=True)
        lengths = lengths.to(self.device, non_blocking=True)
        ac = self.d_autocast_bf16
        self._mark("start")
        recon_pre, recon_post = generator_forward(self.g, self.cfg, real, lengths, self.native_cb2d, self.cb2d_fast_tanh)
        self._mark("generator forward")
        out: Dict[str, Tensor] = {"loss_d": real.new_zeros(())}
        if gan:                                                                     # _train_discriminator
            self.red_d.zero()
            # discriminator gradients are produced on the bands' side streams and on the stream this step runs on (the
            # capture stream when a CUDA graph is being recorded)
            self.red_d.producer_streams = self._band_streams + [torch.cuda.current_stream(self.device)]
            fake = recon_post.detach()
            if self.d_training:
                # two passes, as the reference makes them: in training mode each forward advances the spectral-norm
                # power iteration (first iteration of an epoch only, train.py:417-418 vs :504-506)
                rl, rm, _ = self._patch(real, lengths, True)
                fl, fm, _ = self._patch(fake, lengths, True)
                rl2, rm2, _ = self._multibin(real, lengths, True)
                fl2, fm2, _ = self._multibin(fake, lengths, True)
            else:
                # eval-mode discriminators are per-sample functions: real and fake go through as one batch of 2B
                nb = real.shape[0]
                both, len2 = torch.cat([real, fake], dim=0), torch.cat([lengths, lengths], dim=0)
                lg, mk, _ = self._patch(both, len2, False)
                rl, fl, rm, fm = lg[:nb], lg[nb:], mk[:nb], mk[nb:]
                lg2, mk2, _ = self._multibin(both, len2, False)
                rl2, fl2 = [t_[:nb] for t_ in lg2], [t_[nb:] for t_ in lg2]
                rm2, fm2 = [t_[:nb] for t_ in mk2], [t_[nb:] for t_ in mk2]
            loss_d = self.lecam.d_loss(rl, fl, rm, fm)
            loss_d = loss_d + sum(self.lecam.d_loss(r, f, rm2[0], fm2[0]) for r, f in zip(rl2, fl2)) / len(rl2)
            self._mark("D step: discriminators forward (real + fake) + losses")
            loss_d.backward()
            self.red_d.finish()
            self._mark("D step: backward")
            if clip:
                torch.nn.utils.clip_grad_norm_(self.d_params(), clip)
            self.opt_d.step()
            self._mark("D step: clip + Adam")
            out["loss_d"] = loss_d.detach()
        # _train_generator
        self.red_g.zero()
        self.d_training = False                                                     # train.py:417-418
        d_all = self.d_params_all()
        for p in d_all:                                                             # D is only a loss here: skip its weight gradients
            p.requires_grad_(False)
        try:
            mel = lambda a, g: masked_mel_loss(a, real, lengths, g)
            loss_recon_pre = mel(recon_pre, 1) + 0.25 * mel(recon_pre, 16)
            loss_recon_post = mel(recon_post, 1) + 0.25 * mel(recon_post, 16)
            loss_gan = real.new_zeros(())
            loss_fm = real.new_zeros(())
            gl_lambda = fm_lambda = 0.0
            if gan:
                gl, gm, gf = self._patch(recon_post, lengths, False)
                gl2, gm2, gf2 = self._multibin(recon_post, lengths, False)
                loss_gan = 0.5 * (masked_mse(gl, 1.0, gm) + sum(masked_mse(g, 1.0, gm2[0]) for g in gl2) / len(gl2))
                gl_lambda, fm_lambda = lw["Gloss_lambda"], lw["fm_lambda"]
                if use_fm:                                                          # train.py:454-476
                    with torch.no_grad():
                        _, _, rf = self._patch(real, lengths, False)
                        _, _, rf2 = self._multibin(real, lengths, False)
                    fm_d1 = sum(masked_mae(ff, r, m) for (r, m), (ff, _) in zip(rf, gf)) / max(len(rf), 1)
                    fm_mbd = real.new_zeros(())
                    for rfe, gfe in zip(rf2, gf2):                                  # the reference's running division :466-472
                        for (r, m), (ff, _) in zip(rfe, gfe):
                            fm_mbd = fm_mbd + masked_mae(ff, r, m)
                        if len(rfe) > 0:
                            fm_mbd = fm_mbd / len(rfe)
                    loss_fm = 0.5 * (fm_d1 + fm_mbd / max(len(gf2), 1))
            total = (loss_recon_pre * lw.get("recon_lambda_pre", 1.0) + loss_recon_post * lw.get("recon_lambda_post", 2.0)
                     + loss_gan * gl_lambda + loss_fm * fm_lambda)
            self._mark("G step: mel losses + discriminators forward (fake) + GAN / FM losses")
            total.backward()
            self._mark("G step: backward (discriminators dgrad + generator)")
        finally:
            for p in d_all:
                p.requires_grad_(True)
        self.red_g.finish()
        if clip:
            torch.nn.utils.clip_grad_norm_(list(self.g.values()), clip)
        self.opt_g.step()
        self._mark("G step: reduce + clip + Adam")
        out.update(loss_g_total=total.detach(), loss_recon_pre=loss_recon_pre.detach(),
                   loss_recon_post=loss_recon_post.detach(), loss_gan=loss_gan.detach(), loss_fm=loss_fm.detach())
        self.last_recon = (recon_pre.detach(), recon_post.detach())
        return out

    def d_params_all(self) -> List[Tensor]:
        return [v for sd in (self.pd, self.mb) for k, v in sd.items() if not is_disc_buffer(k)]
