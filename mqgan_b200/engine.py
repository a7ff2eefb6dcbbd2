"""Kernel schedule of the PreEncoder re-encode pass on one B200.

``PreEncoderEngine`` packs a reference state-dict once (weight-norm folded, conv
weights in the tcgen05 kernel's K order, bf16 / bf16x3 split, FSQ code table) and
then runs ``encode`` / ``decode`` as a fixed sequence of C-ABI launches on the
caller's CUDA stream.  Op order follows the reference exactly
(preencoder.py:420-504); what differs is layout (channel-last
everywhere, so the reference's permutes vanish) and fusion (bias / mask / APTx /
residual live in the GEMM epilogues; ConvBlock2D's C-fold expansion is never
materialised).

Precision modes
  encoder "f16x2" (default): every encoder GEMM runs as three fp16 products of 2-term operand
      splits (22 significant bits per operand, weights pre-scaled by a power of two, fp32
      accumulate in TMEM) and all element-wise encoder math is fp32 -> indices equal the fp32
      reference except within rounding distance of an FSQ boundary.
  encoder "bf16x3": the same with six bf16 products of 3-term splits (24 bits, twice the MMAs).
  encoder "bf16": single bf16 pass (index agreement rate is reported, not exact).
  decoder "bf16" (default): bf16 operands, fp32 accumulate, bf16 activations between layers (mel error ~1e-3).
  decoder "f16x2": the fp32-grade mode of the decoder and refiner (the reference's decode is fp32, preencoder.py:453-504):
      every GEMM as three fp16 products of 2-term splits like the encoder, activations carried as two fp16 terms
      [h0 | h1] (22 bits) plus fp32 where an element-wise pass or a residual reads them, precise tanh.  Three times
      the MMA work of the bf16 decoder; mel error ~1e-6 relative.
"""
from __future__ import annotations

import os
from typing import Sequence, Dict, List, Optional, Tuple

import torch

from . import ops
from .ops import PackedConv, pack_conv, pack_upconv
from .spec import PreEncoderConfig


def _fold(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """w = g * v / ||v||_2 over all dims but 0 (weight_norm, dim=0)."""
    try:
        return torch._weight_norm(v.float(), g.float(), 0)
    except Exception:  # pragma: no cover - private op missing
        n = v.float().reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
        return v.float() * (g.float() / n)


def folded_weights(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Plain weights from either weight-norm flavour or already-stripped keys."""
    out: Dict[str, torch.Tensor] = {}
    for k, t in sd.items():
        t = t.detach().cpu()
        if k.endswith(".parametrizations.weight.original1"):
            base = k[: -len(".parametrizations.weight.original1")]
            out[base + ".weight"] = _fold(sd[base + ".parametrizations.weight.original0"].detach().cpu(), t)
        elif k.endswith(".weight_v"):
            base = k[: -len(".weight_v")]
            out[base + ".weight"] = _fold(sd[base + ".weight_g"].detach().cpu(), t)
        elif k.endswith("original0") or k.endswith(".weight_g"):
            continue
        else:
            out[k] = t.float()
    return out


def cb2d_exact_g(s: torch.Tensor, wpw, bpw, wout, bout) -> torch.Tensor:
    """g(s) = sum_k wout_k * aptx(wpw_k s + bpw_k; 1, .5) + bout in the dtype of ``s``
    (preencoder.py:288-295 restated for one scalar pixel value)."""
    u = s[:, None] * wpw[None, :] + bpw[None, :]
    return ((1 + torch.tanh(u)) * 0.5 * u) @ wout + bout


def build_cb2d_table(wpw, bpw, wout, bout, s_max: float = 64.0, tol: float = 2.5e-7, max_intervals: int = 1 << 17,
                     device="cpu"):
    """Per-interval cubic fit of g on [-s_max, s_max) (Chebyshev nodes, float64), refined
    until the float32-evaluated cubic matches the float64 exact sum to ``tol`` * max(1, |g|)
    (about two fp32 ulps) at interior test points.  Load-time work, done with torch float64
    on ``device``.  Returns (coeffs float32 [n,4], offset, inv_h, max_err) or None."""
    wpw, bpw, wout = wpw.double().to(device), bpw.double().to(device), wout.double().to(device)
    nodes = 0.5 - 0.5 * torch.cos((2 * torch.arange(4, dtype=torch.float64, device=device) + 1) * torch.pi / 8)
    V = torch.stack([nodes ** j for j in range(4)], dim=1)                                          # (4 nodes, 4 coeffs)
    Vinv = torch.linalg.inv(V)
    tests = torch.tensor([0.0, 0.21, 0.5, 0.79, 0.999], dtype=torch.float64, device=device)
    inv_h = 32.0
    while True:
        n = int(2 * s_max * inv_h)
        if n > max_intervals:
            return None
        left = (torch.arange(n, dtype=torch.float64, device=device) - n // 2) / inv_h
        pts = (left[:, None] + nodes[None, :] / inv_h).reshape(-1)
        g = torch.cat([cb2d_exact_g(c, wpw, bpw, wout, bout) for c in pts.split(1 << 14)]).reshape(n, 4)
        coef = (g @ Vinv.T).float()                                                                 # (n, 4)
        tp = (left[:, None] + tests[None, :] / inv_h).reshape(-1)
        ref = torch.cat([cb2d_exact_g(c, wpw, bpw, wout, bout) for c in tp.split(1 << 14)]).reshape(n, -1)
        t32 = tests.float()[None, :]
        c = coef
        approx = torch.addcmul(c[:, 0:1], t32, torch.addcmul(c[:, 1:2], t32, torch.addcmul(c[:, 2:3], t32, c[:, 3:4])))
        err = ((approx.double() - ref).abs() / ref.abs().clamp_min(1.0)).max().item()
        if err <= tol:
            return coef.contiguous(), n // 2, inv_h, err
        inv_h *= 2.0


class _CB2D:
    """Packed ConvBlock2D parameters (preencoder.py:251-268)."""

    def __init__(self, w: Dict[str, torch.Tensor], prefix: str, device, use_table: bool = True):
        dw = torch.cat([w[prefix + ".dw.weight"].reshape(25), w[prefix + ".dw.bias"].reshape(1)])
        c = w[prefix + ".pw.weight"].shape[0]
        pw = torch.zeros(c, 4)
        pw[:, 0] = w[prefix + ".pw.weight"].reshape(c)
        pw[:, 1] = w[prefix + ".pw.bias"].reshape(c)
        pw[:, 2] = w[prefix + ".conv_out.weight"].reshape(c)
        self.dw = dw.float().contiguous().to(device)
        self.pw = pw.float().contiguous().to(device)
        self.bout = float(w[prefix + ".conv_out.bias"].reshape(()))
        self.c = c
        self.table, self.table_off, self.table_inv_h, self.table_err = None, 0, 0.0, None
        if use_table and c % 4 == 0:
            t = build_cb2d_table(pw[:, 0], pw[:, 1], pw[:, 2], self.bout, device=device)
            if t is not None:
                self.table = t[0].to(device)
                self.table_off, self.table_inv_h, self.table_err = t[1], t[2], t[3]

    def kwargs(self):
        if self.table is None:
            return {}
        return {"table": self.table, "table_off": self.table_off, "table_inv_h": self.table_inv_h}


class PreEncoderEngine:
    def __init__(self, cfg: PreEncoderConfig, state_dict: Dict[str, torch.Tensor], device="cuda",
                 encoder_precision: str = "f16x2", max_chunk_frames: int = 32768, cb2d_table: bool = True,
                 fuse_upcat: bool = True, max_chunk_frames_enc: int = 262144, fuse_pool: bool = True,
                 decoder_precision: str = "bf16"):
        if encoder_precision not in ("f16x2", "bf16x3", "bf16"):
            raise ValueError("encoder_precision must be 'f16x2', 'bf16x3' or 'bf16'")
        if decoder_precision not in ("bf16", "f16x2"):
            raise ValueError("decoder_precision must be 'bf16' or 'f16x2'")
        self.decoder_precision = decoder_precision
        self.dec_split = decoder_precision == "f16x2"
        dp = "f16x2" if self.dec_split else False      # operand format of every decoder / refiner GEMM
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("PreEncoderEngine needs a CUDA device (no CPU fallback)")
        self.enc_split = encoder_precision != "bf16"
        self.encoder_precision = encoder_precision
        self.max_chunk_frames = int(max_chunk_frames)
        self._copy_stream = None
        self.group_cost_frames = 2048            # decode(lengths_host=...): fixed cost of one more refiner group, in frames
        # the encoder's live set is ~22 KB per frame (the refiner's is ~10x that), so it runs in 8x larger
        # utterance chunks: fewer, fuller waves of GEMM tiles and 8x fewer launches of the CBAM reductions
        self.max_chunk_frames_enc = int(max_chunk_frames_enc)
        self.fsq = ops.fsq_params(cfg.fsq_levels)
        w = folded_weights(state_dict)
        dev = self.device
        sp = encoder_precision        # operand format of every encoder GEMM (ops.pack_conv ``split``)

        def f32(name):
            return w[name].float().contiguous().to(dev)

        # ---------------- encoder ----------------
        self.proj = pack_conv(w["proj.weight"], w["proj.bias"], "linear", sp).to(dev)
        self.pre = _CB2D(w, "pre", dev, cb2d_table)
        self.enc = []
        for i, (cin, cout, k) in enumerate(cfg.encoder_layers):
            p = f"encoder_blocks.{i}"
            blk = {
                "cin": cin, "cout": cout,
                "conv1": pack_conv(w[p + ".conv1.weight"], w[p + ".conv1.bias"], "same1d", sp).to(dev),
                "conv2": pack_conv(w[p + ".conv2.weight"], w[p + ".conv2.bias"], "same1d", sp).to(dev),
                "res": None,
                "mlp_w0": f32(p + ".cbam.channel_attention.mlp.0.weight"),
                "mlp_b0": f32(p + ".cbam.channel_attention.mlp.0.bias"),
                "mlp_w2": f32(p + ".cbam.channel_attention.mlp.2.weight"),
                "mlp_b2": f32(p + ".cbam.channel_attention.mlp.2.bias"),
                "sam_w": w[p + ".cbam.spatial_attention.conv.weight"].reshape(14).float().contiguous().to(dev),
                "beta": float(w[p + ".relu.beta"]), "gamma": float(w[p + ".relu.gamma"]),
            }
            if (p + ".residual.weight") in w:
                blk["res"] = pack_conv(w[p + ".residual.weight"].squeeze(-1), w[p + ".residual.bias"],
                                       "linear", sp).to(dev)
            self.enc.append(blk)
        self.qin_w = f32("q_in_proj.weight")
        self.qin_b = f32("q_in_proj.bias")

        # ---------------- decoder ----------------
        # K8: indices_to_codes + q_out_proj == gather from q_out_proj(implicit_codebook)
        n = cfg.codebook_size
        lv = torch.tensor(cfg.fsq_levels, dtype=torch.int64)
        basis = torch.cumprod(torch.tensor([1] + list(cfg.fsq_levels[:-1])), dim=0)
        digits = (torch.arange(n)[:, None] // basis) % lv                 # quantizer.py:186
        hw = lv // 2
        codes = ((digits - hw) / hw).float()                              # quantizer.py:170
        table = torch.nn.functional.linear(codes, w["q_out_proj.weight"].float(), w["q_out_proj.bias"].float())
        self.code_table = table.contiguous().to(dev)
        self.dec = []
        for i, (cin, cout, k) in enumerate(cfg.decoder_layers):
            p = f"decoder_blocks.{i}"
            blk = {
                "cin": cin, "cout": cout,
                "conv1": pack_conv(w[p + ".conv1.weight"], w[p + ".conv1.bias"], "causal1d", dp).to(dev),
                "conv2": pack_conv(w[p + ".conv2.weight"], w[p + ".conv2.bias"], "causal1d", dp).to(dev),
                "res": None,
                "beta": float(w[p + ".relu.beta"]), "gamma": float(w[p + ".relu.gamma"]),
            }
            if (p + ".residual.weight") in w:
                blk["res"] = pack_conv(w[p + ".residual.weight"].squeeze(-1), w[p + ".residual.bias"],
                                       "linear", dp).to(dev)
            self.dec.append(blk)
        self.post = _CB2D(w, "post", dev, cb2d_table)
        self.out_proj = pack_conv(w["out_proj.weight"], w["out_proj.bias"], "linear", dp).to(dev)
        self.hidden_proj = pack_conv(w["hidden_proj.weight"], w["hidden_proj.bias"], "linear", dp).to(dev)

        # ---------------- refiner ----------------
        chs = cfg.refiner_channels
        d = cfg.refiner_depth
        self.stem_w = w["refiner.pre.conv1.weight"].reshape(chs[0], 9).float().contiguous().to(dev)
        self.stem_b = f32("refiner.pre.conv1.bias")

        def cb(prefix, first=True):
            out = {}
            if first:
                out["conv1"] = pack_conv(w[prefix + ".conv1.weight"], w[prefix + ".conv1.bias"], "conv2d3", dp).to(dev)
            out["conv2"] = pack_conv(w[prefix + ".conv2.weight"], w[prefix + ".conv2.bias"], "conv2d3", dp).to(dev)
            return out

        self.ref_pre = cb("refiner.pre", first=False)
        self.ref_downs = [cb(f"refiner.downs.{i}.conv") for i in range(d)]
        self.ref_mid = cb("refiner.mid")
        self.ref_ups = [cb(f"refiner.ups.{i}.conv") for i in range(d)]
        # the fused upsample-concat conv and the fused pool epilogue exist for bf16 operands only
        self.fuse_upcat = bool(fuse_upcat) and not self.dec_split
        self.fuse_pool = bool(fuse_pool) and not self.dec_split    # DownBlock's AvgPool2d written by the producing conv's epilogue
        for i in range(d if self.fuse_upcat else 0):          # fused nearest-upsample + concat variant of ups[i].conv1
            pfx = f"refiner.ups.{i}.conv.conv1"
            self.ref_ups[i]["conv1_up"] = pack_upconv(w[pfx + ".weight"], w[pfx + ".bias"], chs[d - i], chs[d - i - 1]).to(dev)
        # refiner.post: (1, C, 3, 3) -> (9, C) with tap = 3*(dt+1) + (df+1), run as a 1x1 GEMM C -> 9
        # ... declared 12 wide (three zero rows): the planes buffer has 12 floats per pixel anyway, and a width that is a
        # multiple of four lets the epilogue write it with 16-byte stores
        w9 = torch.zeros(12, chs[0])
        w9[:9] = w["refiner.post.weight"].reshape(chs[0], 9).t()
        self.tail = pack_conv(w9, None, "linear", dp).to(dev)
        # ... or as the 3x3 convolution it is, one output channel on the CTA-pair halo kernel (N tile of 32, weights
        # resident): the tail then reads 4 bytes per pixel instead of nine 4-byte taps out of a 48-byte record
        self.tail3 = pack_conv(w["refiner.post.weight"], None, "conv2d3", dp).to(dev)
        # ... or as its three ROW sums: output channel dt + 1 = the taps of kernel row dt applied along F (three taps in
        # the K loop instead of nine; N tiles of 32 cost the tensor pipe the same whether one or four columns are live),
        # which the tail adds down T.  16 bytes per pixel between the passes.
        wr = torch.zeros(4, chs[0], 3)
        wr[:3] = w["refiner.post.weight"][0].permute(1, 0, 2)            # (dt, C, df)
        self.tail_rows = pack_conv(wr, None, "taps2d", dp, taps=([0, 0, 0], [-1, 0, 1])).to(dev)
        self.post_mode = os.environ.get("MQ_POST_MODE", "rows")           # rows | direct | planes
        self.tail_b = float(w["refiner.post.bias"].reshape(()))
        self.reproj_t = w["refiner.reproj.weight"].t().float().contiguous().to(dev)       # (F, M)

    # ------------------------------------------------------------------
    def _chunks(self, B: int, T: int, frames: Optional[int] = None):
        per = max(1, (frames or self.max_chunk_frames) // max(T, 1))
        for b0 in range(0, B, per):
            yield b0, min(B, b0 + per)

    def _length_groups(self, lengths: Sequence[int], T: int):
        """Partition a ragged batch into length-sorted groups [(member indices, T_group)], T_group = the group's longest
        utterance rounded up to 8 frames plus one coarse row, when that saves at least 10 % of the padded frames; None otherwise.  Among
        1, 2, 4, 8, ... equal-count groups the one minimising (frames computed + a fixed cost per group) wins, and no
        group exceeds ``max_chunk_frames``."""
        n = len(lengths)
        if n < 2:
            return None
        lens = [max(1, min(int(l), T)) for l in lengths]
        mult = 1 << self.cfg.refiner_depth
        order = sorted(range(n), key=lambda i: (lens[i], i))
        per_group_cost = self.group_cost_frames                        # launch tails of ~70 kernels per group
        best = None
        k = 1
        while k <= n:
            size = -(-n // k)
            parts = [order[i:i + size] for i in range(0, n, size)]
            groups, cost = [], 0
            for part in parts:
                # one whole coarse (1/8-resolution) row of padding must follow the longest utterance: ConvBlock does not
                # mask between its two convolutions (preencoder.py:97-98), so the first padded row after an utterance
                # carries aptx(bias + spill) into conv2 - it has to exist, as it does in the padded batch, not be
                # replaced by the image border's zero padding
                Tg = min(T, (-(-max(lens[i] for i in part) // mult) + 1) * mult)
                cap = max(1, self.max_chunk_frames // Tg)              # split further if the group is too large
                for j in range(0, len(part), cap):
                    groups.append((sorted(part[j:j + cap]), Tg))
                    cost += len(part[j:j + cap]) * Tg + per_group_cost
            if best is None or cost < best[0]:
                best = (cost, groups)
            k *= 2
        baseline = n * T + per_group_cost * -(-n * T // self.max_chunk_frames)
        if best[0] > 0.9 * baseline:
            return None
        return best[1]

    @staticmethod
    def _mask_u8(mask: Optional[torch.Tensor], B: int, T: int, device) -> Optional[torch.Tensor]:
        """(B,1,T) / (B,T) bool or uint8, True = padded -> contiguous uint8 (B,T) on device."""
        if mask is None:
            return None
        m = mask.reshape(B, T)
        if m.dtype != torch.uint8:
            m = m.to(torch.uint8)
        return m.to(device).contiguous()

    # ------------------------------------------------------------------
    @ops.on_device
    def encode(self, mel: torch.Tensor, mask: Optional[torch.Tensor] = None, return_latents: bool = False,
               taps: Optional[dict] = None):
        """mel (B,T,n_mels) fp32 on device; mask (B,1,T)|(B,T), True = padded -> (B,T) int64."""
        if mel.dim() != 3 or mel.shape[2] != self.cfg.mel_channels:
            raise ValueError(f"mel must be (B, T, {self.cfg.mel_channels}), got {tuple(mel.shape)}")
        B, T, _ = mel.shape
        mel = mel.to(self.device, torch.float32).contiguous()
        m8 = self._mask_u8(mask, B, T, self.device)
        idx = torch.empty(B, T, dtype=torch.int64, device=self.device)
        zs = torch.empty(B, T, self.cfg.quantizer_dim, dtype=torch.float32, device=self.device) if return_latents else None
        for b0, b1 in self._chunks(B, T, self.max_chunk_frames_enc):
            i, z = self._encode_chunk(mel[b0:b1], None if m8 is None else m8[b0:b1], return_latents, taps)
            idx[b0:b1] = i
            if return_latents:
                zs[b0:b1] = z
        return (idx, zs) if return_latents else idx

    def _encode_chunk(self, mel, m8, want_z, taps):
        cfg, dev, sp = self.cfg, self.device, self.enc_split
        B, T, M = mel.shape
        rows = B * T
        nt = ops.SPLIT_TERMS[self.encoder_precision]
        op_dt = torch.float16 if self.encoder_precision == "f16x2" else torch.bfloat16

        def act_buf(c):
            return torch.empty(rows, nt * c, dtype=op_dt, device=dev)

        def kw_out(buf):
            return {"out_split": buf} if sp else {"out_bf16": buf}

        a0 = ops.split_bf16(mel.reshape(rows, M), nt)
        h = torch.empty(rows, cfg.c0, dtype=torch.float32, device=dev)
        ops.conv_gemm(a0, self.proj, B, T, 1, out_f32=h, tag="enc.proj")                                   # preencoder.py:433
        x32 = torch.empty(rows, cfg.c0, dtype=torch.float32, device=dev)
        xs = act_buf(cfg.c0)
        ops.convblock2d(h, B, T, cfg.c0, self.pre.dw, self.pre.pw, self.pre.bout, m8, False,
                        out_f32=x32, **kw_out(xs), **self.pre.kwargs())                                          # :440
        if taps is not None:
            taps["proj"], taps["pre"] = h, x32
        for i, blk in enumerate(self.enc):                                                  # :443-444
            cout = blk["cout"]
            o1 = act_buf(cout)
            ops.conv_gemm(xs, blk["conv1"], B, T, 1, row_mask=m8, mask_pre=m8 is not None, act=True,
                          beta=blk["beta"], gamma=blk["gamma"], fast_tanh=False, tag=f"enc{i}.conv1", **kw_out(o1))
            o = torch.empty(rows, cout, dtype=torch.float32, device=dev)
            ops.conv_gemm(o1, blk["conv2"], B, T, 1, out_f32=o, tag=f"enc{i}.conv2")
            if blk["res"] is not None:
                r = torch.empty(rows, cout, dtype=torch.float32, device=dev)
                ops.conv_gemm(xs, blk["res"], B, T, 1, out_f32=r, tag=f"enc{i}.res")
            else:
                r = x32
            gate = ops.cam_gate(o, m8, B, T, cout, blk["mlp_w0"], blk["mlp_b0"], blk["mlp_w2"], blk["mlp_b2"])
            y32 = torch.empty(rows, cout, dtype=torch.float32, device=dev)
            last = i == len(self.enc) - 1
            ys = None if last else act_buf(cout)
            ops.cbam_apply(o, gate, r, m8, B, T, cout, blk["sam_w"], blk["beta"], blk["gamma"],
                           out_f32=y32, **({} if last else kw_out(ys)))
            x32, xs = y32, ys
            if taps is not None:
                taps[f"enc{i}"] = y32
        out = ops.qin_fsq(x32, self.qin_w, self.qin_b, self.fsq, want_z=want_z)            # :448-451
        if want_z:
            return out[0].view(B, T), out[1].view(B, T, -1)
        return out.view(B, T), None

    # ------------------------------------------------------------------
    @ops.on_device
    def decode(self, indices: torch.Tensor, mask: Optional[torch.Tensor] = None, return_hidden: bool = False,
               taps: Optional[dict] = None, return_recon: bool = False, host_out: Optional[torch.Tensor] = None,
               lengths_host: Optional[Sequence[int]] = None):
        """indices (B,T) int -> x_post (B,T,n_mels) fp32 [, decoder_out (B,C0,T)] [, x_recon (B,T,n_mels)].

        ``host_out``: optional pinned host tensor (B,T,n_mels) fp32.  Each refiner chunk's result is copied to it on a
        side stream as soon as the chunk is done, so the device-to-host transfer of the re-encoded mels overlaps the
        remaining chunks' compute; the caller's stream waits for the copies before ``decode`` returns.

        ``lengths_host``: the utterance lengths as host integers (must describe the same padding as ``mask``).  The
        decoder and refiner are padding-invariant (each utterance's output depends on its own frames
        only), so a ragged batch is then run through the refiner - 95 % of the FLOPs - in length-sorted groups cut to
        their own longest utterance instead of the batch's: same numbers, less padding to compute."""
        if indices.dim() != 2:
            raise ValueError(f"indices must be (B, T), got {tuple(indices.shape)}")
        B, T = indices.shape
        idx = indices.to(self.device, torch.int64).contiguous()
        m8 = self._mask_u8(mask, B, T, self.device)
        out = torch.empty(B, T, self.cfg.mel_channels, dtype=torch.float32, device=self.device)
        hid = torch.empty(B, T, self.cfg.c0, dtype=torch.float32, device=self.device) if return_hidden else None
        recon = torch.empty_like(out) if return_recon else None
        bad = torch.zeros(1, dtype=torch.int32, device=self.device)      # out-of-range index flag, read once below
        copy_stream = None
        if host_out is not None:
            if host_out.shape != out.shape or host_out.dtype != torch.float32 or not host_out.is_pinned():
                raise ValueError("host_out must be a pinned float32 host tensor of shape (B, T, n_mels)")
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            copy_stream = self._copy_stream
            main = torch.cuda.current_stream(self.device)
        # The 1-D decoder (live set ~10 KB per frame) runs in the large encoder-size chunks; only the
        # refiner (~0.2 MB per frame) is cut into the small ones.
        for c0, c1 in self._chunks(B, T, self.max_chunk_frames_enc):
            mc = None if m8 is None else m8[c0:c1]
            h, R = (self._decode_1d_split if self.dec_split else self._decode_1d)(idx[c0:c1], mc, return_hidden, taps, bad)
            if return_hidden:
                hid[c0:c1] = h
            if return_recon:
                recon[c0:c1] = R.view(c1 - c0, T, -1)[..., : self.cfg.mel_channels]
            Rv = R.view(c1 - c0, T, -1)
            groups = self._length_groups(lengths_host[c0:c1], T) if (lengths_host is not None and taps is None
                                                                    and mc is not None) else None
            if groups is not None:
                out[c0:c1] = Rv[..., : self.cfg.mel_channels]         # padded frames keep x_recon (App. B8)
                for members, Tg in groups:
                    sel = torch.tensor(members, dtype=torch.long, device=self.device)
                    Rg = Rv.index_select(0, sel)[:, :Tg].contiguous()
                    mg = mc.index_select(0, sel)[:, :Tg].contiguous()
                    og = torch.empty(len(members), Tg, self.cfg.mel_channels, dtype=torch.float32, device=self.device)
                    self._refiner(Rg.view(len(members) * Tg, -1), mg, len(members), Tg, og, None)
                    out[c0:c1][sel, :Tg] = og
                if copy_stream is not None:
                    copy_stream.wait_stream(main)
                    with torch.cuda.stream(copy_stream):
                        host_out[c0:c1].copy_(out[c0:c1], non_blocking=True)
                continue
            for b0, b1 in self._chunks(c1 - c0, T):
                self._refiner(Rv[b0:b1].reshape((b1 - b0) * T, -1), None if mc is None else mc[b0:b1], b1 - b0, T,
                              out[c0 + b0:c0 + b1], taps)                                    # :496-499
                if copy_stream is not None:
                    copy_stream.wait_stream(main)
                    with torch.cuda.stream(copy_stream):
                        host_out[c0 + b0:c0 + b1].copy_(out[c0 + b0:c0 + b1], non_blocking=True)
        if copy_stream is not None:
            main.wait_stream(copy_stream)
        # one host sync per decode call (not per chunk: a sync drains the launch queue and idles the GPU)
        if int(bad.item()) != 0:
            raise IndexError("decode: index outside [0, codebook_size)")
        res = [out]
        if return_hidden:
            res.append(hid.permute(0, 2, 1))          # reference layout (B, C0, T), preencoder.py:480
        if return_recon:
            res.append(recon)
        return res[0] if len(res) == 1 else tuple(res)

    def _decode_1d(self, idx, m8, want_hidden, taps, bad):
        """code gather -> causal decoder blocks -> post -> out_proj / hidden_proj: returns (decoder_out | None,
        R = cat[x_recon, hidden] (rows, F) fp32), preencoder.py:464-492."""
        cfg, dev = self.cfg, self.device
        B, T = idx.shape
        rows = B * T
        x, _ = ops.code_gather(idx.reshape(rows), self.code_table, bad=bad)                # preencoder.py:464-466
        for i, blk in enumerate(self.dec):                                                  # :476-477
            cout = blk["cout"]
            o1 = torch.empty(rows, cout, dtype=torch.bfloat16, device=dev)
            ops.conv_gemm(x, blk["conv1"], B, T, 1, row_mask=m8, mask_pre=m8 is not None, act=True,
                          beta=blk["beta"], gamma=blk["gamma"], out_bf16=o1, tag=f"dec{i}.conv1")
            if blk["res"] is not None:
                r = torch.empty(rows, cout, dtype=torch.bfloat16, device=dev)
                ops.conv_gemm(x, blk["res"], B, T, 1, out_bf16=r, tag=f"dec{i}.res")
            else:
                r = x
            y = torch.empty(rows, cout, dtype=torch.bfloat16, device=dev)
            ops.conv_gemm(o1, blk["conv2"], B, T, 1, row_mask=m8, mask_pre=m8 is not None, act=True,
                          beta=blk["beta"], gamma=blk["gamma"], res=r, res_mode=1, out_bf16=y, tag=f"dec{i}.conv2")
            x = y
            if taps is not None:
                taps[f"dec{i}"] = y
        dec_out = x
        pz = torch.empty(rows, cfg.c0, dtype=torch.bfloat16, device=dev)
        ops.convblock2d(dec_out, B, T, cfg.c0, self.post.dw, self.post.pw, self.post.bout, m8, True,
                        out_bf16=pz, **self.post.kwargs())                                                        # :482
        F, M = cfg.refiner_width, cfg.mel_channels
        R = torch.empty(rows, F, dtype=torch.float32, device=dev)
        ops.conv_gemm(pz, self.out_proj, B, T, 1, out_f32=R, f32_coff=0, tag="dec.out_proj")                    # :486
        ops.conv_gemm(dec_out, self.hidden_proj, B, T, 1, out_f32=R, f32_coff=M, tag="dec.hidden_proj")            # :490-492
        if taps is not None:
            taps["refiner_in"] = R
        return (dec_out.view(B, T, cfg.c0).float() if want_hidden else None), R

    def _decode_1d_split(self, idx, m8, want_hidden, taps, bad):
        """_decode_1d in the fp32-grade mode: activations as two fp16 terms + fp32, precise tanh."""
        cfg, dev = self.cfg, self.device
        B, T = idx.shape
        rows = B * T
        _, x32 = ops.code_gather(idx.reshape(rows), self.code_table, bf16=False, f32=True, bad=bad)
        xs = ops.split_bf16(x32, 2)
        for i, blk in enumerate(self.dec):
            cout = blk["cout"]
            o1 = torch.empty(rows, 2 * cout, dtype=torch.float16, device=dev)
            ops.conv_gemm(xs, blk["conv1"], B, T, 1, row_mask=m8, mask_pre=m8 is not None, act=True, beta=blk["beta"],
                          gamma=blk["gamma"], fast_tanh=False, out_split=o1, tag=f"dec{i}.conv1")
            if blk["res"] is not None:
                r32 = torch.empty(rows, cout, dtype=torch.float32, device=dev)
                ops.conv_gemm(xs, blk["res"], B, T, 1, out_f32=r32, tag=f"dec{i}.res")
            else:
                r32 = x32
            y32 = torch.empty(rows, cout, dtype=torch.float32, device=dev)
            ys = torch.empty(rows, 2 * cout, dtype=torch.float16, device=dev)
            ops.conv_gemm(o1, blk["conv2"], B, T, 1, row_mask=m8, mask_pre=m8 is not None, act=True, beta=blk["beta"],
                          gamma=blk["gamma"], fast_tanh=False, res=r32, res_mode=1, out_f32=y32, out_split=ys,
                          tag=f"dec{i}.conv2")
            x32, xs = y32, ys
            if taps is not None:
                taps[f"dec{i}"] = y32
        pz = torch.empty(rows, 2 * cfg.c0, dtype=torch.float16, device=dev)
        ops.convblock2d(x32, B, T, cfg.c0, self.post.dw, self.post.pw, self.post.bout, m8, False,
                        out_split=pz, **self.post.kwargs())
        F, M = cfg.refiner_width, cfg.mel_channels
        R = torch.empty(rows, F, dtype=torch.float32, device=dev)
        ops.conv_gemm(pz, self.out_proj, B, T, 1, out_f32=R, f32_coff=0, tag="dec.out_proj")
        ops.conv_gemm(xs, self.hidden_proj, B, T, 1, out_f32=R, f32_coff=M, tag="dec.hidden_proj")
        if taps is not None:
            taps["refiner_in"] = R
        return (x32.view(B, T, cfg.c0) if want_hidden else None), R

    def _refiner_split(self, R, m8, B, T, out, taps):
        """_refiner in the fp32-grade mode (preencoder.py:169-202): two-term fp16 activations, three fp16 products per
        GEMM on the CTA-pair kernel, AvgPool / upsample-concat as separate split-aware passes."""
        cfg, dev = self.cfg, self.device
        d = cfg.refiner_depth
        chs = cfg.refiner_channels
        F = cfg.refiner_width
        T8, down, up = ops.refiner_masks(m8, B, T, d, dev)
        H = [T8 >> l for l in range(d + 1)]

        def buf(l, c):
            return torch.empty(B, H[l], F, 2 * c, dtype=torch.float16, device=dev)

        def convblock(x, blk, l, mask, cout, first=True, tag="", x32=None, want_f32=False):
            """ConvBlock (preencoder.py:86-102); x32 = fp32 copy of the input when the block has the +x skip."""
            if first:
                t = buf(l, cout)
                ops.conv_gemm(x, blk["conv1"], B, H[l], F, act=True, fast_tanh=False, out_split=t, tag=tag + ".conv1")
            else:
                t = x
            y = buf(l, cout)
            y32 = torch.empty(B, H[l], F, cout, dtype=torch.float32, device=dev) if want_f32 else None
            ops.conv_gemm(t, blk["conv2"], B, H[l], F, act=True, fast_tanh=False, row_mask=mask, mask_post=True,
                          res=x32, res_mode=2 if x32 is not None else 0, out_split=y, out_f32=y32, tag=tag + ".conv2")
            return (y, y32) if want_f32 else y

        s1 = ops.refiner_stem_split(R, m8, B, T, T8, F, chs[0], self.stem_w, self.stem_b)
        x = convblock(s1, self.ref_pre, 0, down[0], chs[0], first=False, tag="ref.pre")
        skips = []
        x32 = None
        for i in range(d):
            skips.append(x)
            p = ops.avgpool_mask_split(x, down[i + 1], B, H[i], F, chs[i])
            if i + 1 < d:
                x = convblock(p, self.ref_downs[i], i + 1, down[i + 1], chs[i + 1], tag=f"ref.down{i}")
            else:       # the deepest block's output is mid's residual: also kept in fp32
                x, x32 = convblock(p, self.ref_downs[i], i + 1, down[i + 1], chs[i + 1], tag=f"ref.down{i}", want_f32=True)
            if taps is not None:
                taps[f"refiner.downs.{i}"] = x
        x = convblock(x, self.ref_mid, d, down[d], chs[d], tag="ref.mid", x32=x32)
        if taps is not None:
            taps["refiner.mid"] = x
        for i in range(d):
            l = d - 1 - i
            skip = skips.pop()
            u = ops.upcat_mask_split(x, skip, up[l], B, H[l], F, chs[l + 1], chs[l])
            x = convblock(u, self.ref_ups[i], l, up[l], chs[l], tag=f"ref.up{i}")
            if taps is not None:
                taps[f"refiner.ups.{i}"] = x
        self._post_tail(x, m8, B, T, T8, F, R, out)

    def _post_tail(self, x, m8, B, T, T8, F, R, out):
        """refiner.post (C -> 1, 3x3; preencoder.py:191) + crop / mask / reproj / + x_recon (:192-200, :499)."""
        dev = x.device
        mode = self.post_mode if (ops.PAIR_DEFAULT and F >= 8) else "planes"
        # rows / direct: always the CTA-pair kernel, whatever T8 - a length group of a few frames must produce the bits
        # the padded batch produces (one K order per output pixel), and F is a property of the model, not of the batch
        if mode == "rows":
            tp = torch.empty(B, T8, F, 4, dtype=torch.float32, device=dev)
            ops.conv_gemm(x, self.tail_rows, B, T8, F, out_f32=tp, pair=True, tag="ref.post")
        elif mode == "direct":
            tp = torch.empty(B, T8, F, 1, dtype=torch.float32, device=dev)
            ops.conv_gemm(x, self.tail3, B, T8, F, out_f32=tp, pair=True, tag="ref.post")
        else:
            tp = torch.empty(B, T8, F, 12, dtype=torch.float32, device=dev)
            ops.conv_gemm(x, self.tail, B, T8, F, out_f32=tp, tag="ref.post")               # as 9 tap planes
        ops.refiner_tail(tp, m8, B, T, T8, F, self.tail_b, self.reproj_t, self.cfg.mel_channels, R, out=out)

    def _refiner(self, R, m8, B, T, out, taps):
        if self.dec_split:
            return self._refiner_split(R, m8, B, T, out, taps)
        cfg, dev = self.cfg, self.device
        d = cfg.refiner_depth
        chs = cfg.refiner_channels
        F = cfg.refiner_width
        T8, down, up = ops.refiner_masks(m8, B, T, d, dev)
        H = [T8 >> l for l in range(d + 1)]

        def convblock(x, blk, l, mask, cin, cout, first=True, tag="", pool=False):
            """ConvBlock (preencoder.py:86-102).  pool=True also returns AvgPool2d((2,1)) of the output,
            filled by the max-pooled mask, written by conv2's epilogue (DownBlock, :111-114)."""
            if first:
                t = torch.empty(B, H[l], F, cout, dtype=torch.bfloat16, device=dev)
                ops.conv_gemm(x, blk["conv1"], B, H[l], F, act=True, out_bf16=t, tag=tag + ".conv1")   # preencoder.py:97
            else:
                t = x
            y = torch.empty(B, H[l], F, cout, dtype=torch.bfloat16, device=dev)
            match = first and cin == cout
            # needs the 8-pixel-wide sub-tiles of the halo / CTA-pair main loops
            fused_pool = (pool and self.fuse_pool and F >= 8 and cout % 32 == 0
                          and (ops.HALO_DEFAULT or (ops.PAIR_DEFAULT and H[l] >= 32)))
            yp = torch.empty(B, H[l] // 2, F, cout, dtype=torch.bfloat16, device=dev) if fused_pool else None
            ops.conv_gemm(t, blk["conv2"], B, H[l], F, act=True, row_mask=mask, mask_post=True,
                          res=x if match else None, res_mode=2 if match else 0, out_bf16=y, out_pool=yp,
                          tag=tag + ".conv2")                                               # :98-101
            if pool and not fused_pool:
                yp = ops.avgpool_mask(y, down[l + 1], B, H[l], F, cout)
            return (y, yp) if pool else y

        s1 = ops.refiner_stem(R, m8, B, T, T8, F, chs[0], self.stem_w, self.stem_b, True)   # :172-175, :97
        x, p = convblock(s1, self.ref_pre, 0, down[0], 1, chs[0], first=False, tag="ref.pre", pool=True)
        skips = []
        for i in range(d):                                                                  # :179-181
            skips.append(x)
            if i + 1 < d:
                x, p = convblock(p, self.ref_downs[i], i + 1, down[i + 1], chs[i], chs[i + 1], tag=f"ref.down{i}",
                                 pool=True)
            else:
                x = convblock(p, self.ref_downs[i], i + 1, down[i + 1], chs[i], chs[i + 1], tag=f"ref.down{i}")
            if taps is not None:
                taps[f"refiner.downs.{i}"] = x
        x = convblock(x, self.ref_mid, d, down[d], chs[d], chs[d], tag="ref.mid")                          # :184
        if taps is not None:
            taps["refiner.mid"] = x
        for i in range(d):                                                                  # :187-189
            l = d - 1 - i
            skip = skips.pop()
            if self.fuse_upcat:
                # upsample + cat + mask folded into conv1's operand loads (no (Cx+Cs)-channel copy)
                if m8 is not None or T8 != T:       # without a mask and without T padding both masks are all-valid
                    ops.zero_rows(skip, up[l], down[l])
                t = torch.empty(B, H[l], F, chs[l], dtype=torch.bfloat16, device=dev)
                ops.conv_gemm(x, self.ref_ups[i]["conv1_up"], B, H[l + 1], F, x2=skip, act=True, out_bf16=t,
                              tag=f"ref.up{i}.conv1")
                x = convblock(t, self.ref_ups[i], l, up[l], 0, chs[l], first=False, tag=f"ref.up{i}")
            else:
                u = ops.upcat_mask(x, skip, up[l], B, H[l], F, chs[l + 1], chs[l])
                x = convblock(u, self.ref_ups[i], l, up[l], chs[l + 1] + chs[l], chs[l], tag=f"ref.up{i}")
            if taps is not None:
                taps[f"refiner.ups.{i}"] = x
        self._post_tail(x, m8, B, T, T8, F, R, out)                                         # :191-200, :499
