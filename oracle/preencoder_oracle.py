"""CPU oracle for the PreEncoder re-encode path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this module, and only as the checker or
the timed CPU baseline.  Nothing under ``mqgan_b200/`` imports it.

It is a functional restatement (state-dict in, tensors out; no nn.Module, no
code from the reference) of the eval-mode algorithm of

    preencoder.py:420-451   PreEncoder.encode
    preencoder.py:453-504   PreEncoder.decode
    preencoder.py:277-301   ConvBlock2D.forward          (`pre`, `post`)
    preencoder.py:86-130    ConvBlock / DownBlock / UpBlock
    preencoder.py:169-202   UNetRefiner.forward, :29-47 pad_to_pow2_4d
    attentions.py:525-551   ResidualBlock1D.forward
    attentions.py:217-273   CAM1D (non-causal branch), :81-132 pools
    attentions.py:310-365   SAM1D, :393-419 CBAM1D, :471-474 CausalConv1da
    attentions.py:34-35     APTx
    quantizer.py:109-114    FSQ.bound, :128-140 quantize, :177-181
                            codes_to_indices, :183-187 indices_to_level_indices

including the reference's observable quirks: the module-level
``masked_fill_`` helper is not in-place, so CAM's max-pool sees padded frames,
SAM masks nothing and CBAM's output is not zeroed at padded frames.

Parity pin: the reference ships no tests / golden vectors for this path
(SURVEY §4, §8c), so this oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF, run in the build container by ``oracle/make_golden.py`` and committed
under ``tests/golden/`` (``tests/test_oracle_golden.py`` re-checks on every run).
Arithmetic below the op level is PyTorch's CPU library (oneDNN/MKL), the same
dependency the reference calls; ``dtype=torch.float64`` gives the margin oracle.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------
def fold_weight_norm(g: Tensor, v: Tensor) -> Tensor:
    """w = g * v / ||v||, norm over every dim but 0 (weight_norm dim=0;
    preencoder.py:53,248, attentions.py:469,500-501)."""
    return torch._weight_norm(v, g, 0)


def effective_weights(sd: Dict[str, Tensor], dtype=torch.float32) -> Dict[str, Tensor]:
    """Plain ``<module>.weight`` / ``.bias`` tensors with both weight-norm
    flavours folded."""
    out: Dict[str, Tensor] = {}
    for k, t in sd.items():
        t = t.detach().to("cpu")
        if k.endswith(".parametrizations.weight.original1"):
            base = k[: -len(".parametrizations.weight.original1")]
            g = sd[base + ".parametrizations.weight.original0"].detach().cpu()
            out[base + ".weight"] = fold_weight_norm(g.to(dtype), t.to(dtype))
        elif k.endswith(".weight_v"):
            base = k[: -len(".weight_v")]
            g = sd[base + ".weight_g"].detach().cpu()
            out[base + ".weight"] = fold_weight_norm(g.to(dtype), t.to(dtype))
        elif k.endswith("original0") or k.endswith(".weight_g"):
            continue
        else:
            out[k] = t.to(dtype)
    return out


def sequence_mask(max_length: int, lengths: Tensor) -> Tensor:
    """(B, T) bool, True = padded (preencoder.py:15-24)."""
    return torch.arange(max_length)[None, :] >= lengths.reshape(-1, 1).cpu()


# ----------------------------------------------------------------------------
# element-wise pieces
# ----------------------------------------------------------------------------
def aptx(x: Tensor, beta, gamma) -> Tensor:
    """(1 + tanh(beta x)) * gamma * x   (attentions.py:34-35, alpha = 1)."""
    return (1 + torch.tanh(beta * x)) * gamma * x


# ----------------------------------------------------------------------------
# FSQ (quantizer.py)
# ----------------------------------------------------------------------------
def fsq_constants(levels: Sequence[int], dtype=torch.float32):
    lv = torch.tensor(list(levels), dtype=torch.int32)
    basis = torch.cumprod(torch.tensor([1] + list(levels[:-1])), dim=0).to(torch.int32)
    eps = 1e-3
    half_l = (lv - 1) * (1 + eps) / 2                      # quantizer.py:111 (float32 tensor)
    offset = torch.where(lv % 2 == 0, 0.5, 0.0)            # :112
    shift = (offset / half_l).atanh()                      # :113
    half_w = lv // 2                                       # :132
    return lv, basis, half_l.to(dtype), offset.to(dtype), shift.to(dtype), half_w


def fsq_quantize(z: Tensor, levels: Sequence[int]) -> Tuple[Tensor, Tensor]:
    """z (..., D) -> (codes (..., D) in [-1, 1], indices (...) int64).
    quantizer.py:109-114,137,164-166,177-181; round = half-to-even."""
    lv, basis, half_l, offset, shift, half_w = fsq_constants(levels, z.dtype)
    bounded = (z + shift).tanh() * half_l - offset
    q = bounded.round() / half_w
    zhat = q * half_w + half_w
    idx = (zhat * basis).sum(dim=-1).to(torch.int32)
    return q, idx.long()


def fsq_round_margin(z: Tensor, levels: Sequence[int]) -> Tensor:
    """Distance of the bounded latent to the nearest rounding boundary (k + 0.5),
    minimised over the D dims; used for the margin-aware index gate."""
    lv, basis, half_l, offset, shift, half_w = fsq_constants(levels, z.dtype)
    bounded = (z + shift).tanh() * half_l - offset
    frac = bounded - torch.floor(bounded)
    return (frac - 0.5).abs().min(dim=-1).values


def fsq_indices_to_codes(idx: Tensor, levels: Sequence[int], dtype=torch.float32) -> Tensor:
    """(...) int -> (..., D) codes; quantizer.py:183-187, :168-170, :172-175."""
    lv, basis, _, _, _, half_w = fsq_constants(levels, dtype)
    digits = (idx.unsqueeze(-1) // basis) % lv
    return ((digits - half_w) / half_w).to(dtype)


# ----------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------
def convblock2d(x: Tensor, mask: Tensor, w: Dict[str, Tensor], prefix: str, chunk: int = 64) -> Tensor:
    """``pre`` / ``post`` (preencoder.py:277-301).  x (B, C, T), mask (B, 1, T).

    The reference expands to (B, C, C, T); the expansion is point-wise, so it is
    evaluated here in T-chunks to bound memory (same arithmetic per element).
    """
    B, C, T = x.shape
    img = x.unsqueeze(1)                                         # (B,1,C,T)
    m4 = mask.unsqueeze(1)                                       # (B,1,1,T)
    s = F.conv2d(img, w[prefix + ".dw.weight"], w[prefix + ".dw.bias"], padding=2)   # :286
    s = s.masked_fill(m4, 0.0)                                                        # :287
    outs = []
    for t0 in range(0, T, chunk):
        sc = s[..., t0:t0 + chunk]
        mc = m4[..., t0:t0 + chunk]
        u = F.conv2d(sc, w[prefix + ".pw.weight"], w[prefix + ".pw.bias"])            # :288 (B,C,C,Tc)
        u = u.masked_fill(mc, 0.0)                                                    # :292
        a = aptx(u, 1, 0.5)                                                           # :293
        o = F.conv2d(a, w[prefix + ".conv_out.weight"], w[prefix + ".conv_out.bias"])  # :295
        outs.append(o)
    return torch.cat(outs, dim=-1).squeeze(1)                                         # :296


def cbam(o: Tensor, mask: Tensor, w: Dict[str, Tensor], prefix: str) -> Tensor:
    """CBAM1D with the reference's effective (quirky) masking.  o (B, C, T)."""
    # CAM1D (attentions.py:248-273): max over ALL t (helper no-op, App. B1);
    # masked mean over valid t (:109-131, real in-place fill).
    mx = o.max(dim=-1).values
    valid = (~mask).to(o.dtype)                                   # (B,1,T)
    sm = (o * valid).sum(dim=-1)
    cnt = valid.sum(dim=-1).clamp(min=1.0)                        # (B,1)
    av = sm / cnt
    p = prefix + ".channel_attention.mlp."

    def mlp(v):
        h = F.relu(F.linear(v, w[p + "0.weight"], w[p + "0.bias"]))
        return F.linear(h, w[p + "2.weight"], w[p + "2.bias"])

    gate = torch.sigmoid(mlp(mx) + mlp(av)).unsqueeze(-1)         # :262-265
    o1 = gate * o                                                 # :268 (trailing mask is a no-op)
    # SAM1D (attentions.py:322-365): unmasked channel max / mean, conv k7, sigmoid
    pm = o1.max(dim=1, keepdim=True).values
    pa = o1.mean(dim=1, keepdim=True)
    logits = F.conv1d(torch.cat((pm, pa), dim=1), w[prefix + ".spatial_attention.conv.weight"],
                      None, padding=3)
    o2 = torch.sigmoid(logits) * o1
    return o2 + o                                                 # :411


def residual_block(x: Tensor, mask: Tensor, w: Dict[str, Tensor], prefix: str, k: int,
                   causal: bool) -> Tensor:
    """ResidualBlock1D.forward (attentions.py:525-551), eval mode."""
    beta = w[prefix + ".relu.beta"]
    gamma = w[prefix + ".relu.gamma"]
    if (prefix + ".residual.weight") in w:
        r = F.conv1d(x, w[prefix + ".residual.weight"], w[prefix + ".residual.bias"])
    else:
        r = x

    def conv(inp, name):
        if causal:                                                # attentions.py:471-474
            inp = F.pad(inp, (k - 1, 0))
            return F.conv1d(inp, w[f"{prefix}.{name}.weight"], w[f"{prefix}.{name}.bias"])
        return F.conv1d(inp, w[f"{prefix}.{name}.weight"], w[f"{prefix}.{name}.bias"],
                        padding=(k - 1) // 2)                     # padding="same", odd k

    o = conv(x, "conv1").masked_fill(mask, 0)                     # :533-537
    o = aptx(o, beta, gamma)                                      # :538
    o = conv(o, "conv2")                                          # :541 (not masked)
    if not causal:
        o = cbam(o, mask, w, prefix + ".cbam")                    # :543-544
    o = (o + r).masked_fill(mask, 0)                              # :545-548
    return aptx(o, beta, gamma)                                   # :549


def refiner_convblock(x: Tensor, m4: Tensor, w: Dict[str, Tensor], prefix: str) -> Tensor:
    """ConvBlock.forward (preencoder.py:95-102). x (B, C, T, F), m4 (B,1,T,1)."""
    x = x.masked_fill(m4, 0.0)
    y = aptx(F.conv2d(x, w[prefix + ".conv1.weight"], w[prefix + ".conv1.bias"], padding=1), 1, 0.5)
    y = aptx(F.conv2d(y, w[prefix + ".conv2.weight"], w[prefix + ".conv2.bias"], padding=1), 1, 0.5)
    if w[prefix + ".conv1.weight"].shape[0] == w[prefix + ".conv1.weight"].shape[1]:
        y = y + x
    return y.masked_fill(m4, 0.0)


def refiner(r_in: Tensor, mask_bt: Tensor, w: Dict[str, Tensor], depth: int,
            taps: Optional[dict] = None) -> Tensor:
    """UNetRefiner.forward (preencoder.py:169-202). r_in (B, 1, T, F); mask (B, T)."""
    B, _, T, Fw = r_in.shape
    mult = 1 << depth
    pad = (mult - (T % mult)) % mult                              # :38
    x = torch.cat((r_in, r_in.new_zeros(B, 1, pad, Fw)), dim=2)
    m = torch.cat((mask_bt, mask_bt.new_ones(B, pad)), dim=1).reshape(B, 1, T + pad, 1)
    skips: List[Tensor] = []
    x = refiner_convblock(x, m, w, "refiner.pre")
    if taps is not None:
        taps["refiner.pre"] = x
    cur = m
    for i in range(depth):                                        # :179-181
        skips.append(x)
        x = F.avg_pool2d(x, kernel_size=(2, 1))                   # :112
        cur = F.max_pool2d(cur.to(x.dtype), kernel_size=(2, 1), stride=(2, 1)).bool()   # :65
        x = refiner_convblock(x, cur, w, f"refiner.downs.{i}.conv")
        if taps is not None:
            taps[f"refiner.downs.{i}"] = x
    x = refiner_convblock(x, cur, w, "refiner.mid")               # :184
    if taps is not None:
        taps["refiner.mid"] = x
    for i in range(depth):                                        # :187-189
        skip = skips.pop()
        x = F.interpolate(x, scale_factor=(2, 1), mode="nearest")            # :124
        cur = F.interpolate(cur.to(x.dtype), scale_factor=(2, 1), mode="nearest").bool()   # :70
        x = torch.cat([x, skip], dim=1)                           # :129 (crop is a no-op: T % 2^depth == 0)
        x = refiner_convblock(x, cur, w, f"refiner.ups.{i}.conv")
        if taps is not None:
            taps[f"refiner.ups.{i}"] = x
    out = F.conv2d(x.masked_fill(cur, 0.0), w["refiner.post.weight"], w["refiner.post.bias"], padding=1)
    out = out.squeeze(1)[:, :T, :]                                # :192-195
    out = out.masked_fill(mask_bt.unsqueeze(-1), 0.0)             # :198
    return F.linear(out, w["refiner.reproj.weight"])              # :200


# ----------------------------------------------------------------------------
# top level
# ----------------------------------------------------------------------------
def _cfg_layers(cfg):
    return cfg.encoder_layers, cfg.decoder_layers


def encode_latents(sd_or_w, cfg, mel: Tensor, mask: Optional[Tensor] = None, dtype=torch.float32,
                   taps: Optional[dict] = None, folded: bool = False) -> Tensor:
    """mel (B, T, n_mels) -> pre-quantiser latents z (B, T, D).  preencoder.py:433-448."""
    w = sd_or_w if folded else effective_weights(sd_or_w, dtype)
    x = F.linear(mel.to(dtype), w["proj.weight"], w["proj.bias"]).permute(0, 2, 1)       # :433-435
    if mask is None:
        mask = torch.zeros((x.size(0), 1, x.size(2)), dtype=torch.bool)                  # :437-438
    if taps is not None:
        taps["proj"] = x
    x = convblock2d(x, mask, w, "pre")                                                   # :440
    if taps is not None:
        taps["pre"] = x
    for i, (_, _, k) in enumerate(cfg.encoder_layers):                                   # :443-444
        x = residual_block(x, mask, w, f"encoder_blocks.{i}", k, causal=False)
        if taps is not None:
            taps[f"enc{i}"] = x
    x = x.permute(0, 2, 1)
    return F.linear(x, w["q_in_proj.weight"], w["q_in_proj.bias"])                       # :448


def encode(sd_or_w, cfg, mel: Tensor, mask: Optional[Tensor] = None, dtype=torch.float32,
           folded: bool = False) -> Tensor:
    """PreEncoder.encode: (B, T, n_mels) [+ (B,1,T) mask] -> (B, T) int64."""
    z = encode_latents(sd_or_w, cfg, mel, mask, dtype, folded=folded)
    return fsq_quantize(z, cfg.fsq_levels)[1]                                            # :450-451


def decode(sd_or_w, cfg, indices: Tensor, mask: Optional[Tensor] = None, dtype=torch.float32,
           taps: Optional[dict] = None, folded: bool = False, return_hidden: bool = False):
    """PreEncoder.decode: (B, T) int -> (B, T, n_mels).  preencoder.py:453-504."""
    w = sd_or_w if folded else effective_weights(sd_or_w, dtype)
    codes = fsq_indices_to_codes(indices.cpu(), cfg.fsq_levels, dtype)                   # :464
    x = F.linear(codes, w["q_out_proj.weight"], w["q_out_proj.bias"]).permute(0, 2, 1)   # :466-469
    if mask is None:
        mask = torch.zeros((x.size(0), 1, x.size(2)), dtype=torch.bool)                  # :471-472
    dec = x
    for i, (_, _, k) in enumerate(cfg.decoder_layers):                                   # :476-477
        dec = residual_block(dec, mask, w, f"decoder_blocks.{i}", k, causal=True)
        if taps is not None:
            taps[f"dec{i}"] = dec
    xr = convblock2d(dec, mask, w, "post").permute(0, 2, 1)                              # :482-484
    x_recon = F.linear(xr, w["out_proj.weight"], w["out_proj.bias"])                     # :486
    hid = F.linear(dec.permute(0, 2, 1), w["hidden_proj.weight"], w["hidden_proj.bias"])  # :490
    r_in = torch.cat([x_recon, hid], dim=2).unsqueeze(1)                                 # :492-493
    if taps is not None:
        taps["x_recon"] = x_recon
        taps["refiner_in"] = r_in
    res = refiner(r_in, mask.squeeze(1), w, cfg.refiner_depth, taps)                     # :496-498
    if taps is not None:
        taps["residual"] = res
    x_post = x_recon + res                                                               # :499
    if return_hidden:
        return x_post, dec
    return x_post


def reencode(sd, cfg, mel: Tensor, lengths: Optional[Tensor] = None, dtype=torch.float32):
    """encode -> decode as the CLIs do it (reencode_spectrograms.py:65-66)."""
    w = effective_weights(sd, dtype)
    mask = None
    if lengths is not None:
        mask = sequence_mask(mel.shape[1], torch.as_tensor(lengths)).unsqueeze(1)
    idx = encode(w, cfg, mel, mask, dtype, folded=True)
    out = decode(w, cfg, idx, mask, dtype, folded=True)
    return idx, out
