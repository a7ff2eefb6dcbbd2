"""Thin torch-tensor wrappers over the C ABI (one function per entry point) plus
the host-side weight packing for the tcgen05 convolution kernel.

PyTorch is plumbing here: it owns device memory and the stream; every byte of
arithmetic on the hot path happens inside libmqgan_b200.so.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ConvParams, Cb2dParams, CbamApplyParams, FsqParams, VqParams

BLOCK_K = 64
HALO_DEFAULT = os.environ.get("MQ_HALO", "1") != "0"   # halo-tile main loop for 3x3 convs
PAIR_DEFAULT = os.environ.get("MQ_PAIR", "1") != "0"   # CTA-pair (cta_group::2) main loop for the refiner's 3x3 convs
PAIR_MIN_BN = int(os.environ.get("MQ_PAIR_MIN_BN", "0"))
PAIR_1D = os.environ.get("MQ_PAIR_1D", "1") != "0"     # row-halo CTA-pair loop for 1-D convolutions
PAIR_1D_MIN_BN = int(os.environ.get("MQ_PAIR_1D_MIN_BN", "128"))
MSUB_OVERRIDE = int(os.environ.get("MQ_MSUB", "0"))            # experiment knobs, read once at import
MSUB_PAIR_OVERRIDE = int(os.environ.get("MQ_MSUB_PAIR", "0"))
MSUB_PAIR_WIDE = int(os.environ.get("MQ_MSUB_PAIR_WIDE", "1"))   # sub-tiles per CTA for bn = 256 pair layers (2 = one TMEM buffer)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def on_device(fn):
    """Method decorator: run with ``self.device`` as the current CUDA device.  Every launch in this module goes to the
    CURRENT device's current stream, so an engine living on cuda:1 must not be driven while cuda:0 is current (the
    reference accepts ``get_pre_encoder(path, 'cuda:1')`` without a ``set_device``)."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *args, **kwargs):
        dev = torch.device(self.device)
        if dev.type != "cuda" or (dev.index is not None and dev.index == torch.cuda.current_device()):
            return fn(self, *args, **kwargs)
        with torch.cuda.device(dev):
            return fn(self, *args, **kwargs)
    return wrapped


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def _chk(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (this path has no CPU fallback)")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    return t


# ---------------------------------------------------------------------------
# weight packing for mq_conv_gemm
# ---------------------------------------------------------------------------
@dataclass
class PackedConv:
    wpack: torch.Tensor            # bf16 [cout_pad, K]
    bias: Optional[torch.Tensor]   # fp32 [cout]
    cin: int
    cout: int
    cout_pad: int
    bn: int
    taps: int
    nseg: int
    kchunks: int
    tap_dh: List[int]
    tap_dw: List[int]
    a_coff: List[int]
    split: bool
    # fused upsample + concat (UpBlock) extras
    up_taps: int = 0
    kchunks2: int = 0
    cin2: int = 0
    tap_dh_odd: Optional[List[int]] = None
    # operand format: "bf16" (single pass), "bf16x3" (six segments), "f16x2" (three fp16 segments)
    mode: str = "bf16"
    acc_scale: float = 1.0         # epilogue multiplies the accumulator by this (f16x2 weight pre-scale undone)

    def to(self, device):
        self.wpack = self.wpack.to(device)
        if self.bias is not None:
            self.bias = self.bias.to(device)
        return self


def split3_bf16(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """fp32 -> three bf16 terms with x0 + x1 + x2 == x (host-side weight prep)."""
    x = x.float()
    x0 = x.to(torch.bfloat16)
    r = x - x0.float()
    x1 = r.to(torch.bfloat16)
    r = r - x1.float()
    x2 = r.to(torch.bfloat16)
    return x0, x1, x2


def split2_f16(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 -> two fp16 terms with h0 + h1 == x to 22 significant bits (host-side weight prep)."""
    x = x.float().clamp(-65504.0, 65504.0)
    h0 = x.to(torch.float16)
    h1 = (x - h0.float()).to(torch.float16)
    return h0, h1


def split_mode(split) -> str:
    """Normalise the ``split`` argument of pack_conv: False/None -> "bf16", True -> "bf16x3"."""
    if split is None or split is False:
        return "bf16"
    if split is True:
        return "bf16x3"
    if split in ("bf16", "bf16x3", "f16x2"):
        return split
    raise ValueError(f"unknown operand split {split!r}")


SPLIT_TERMS = {"bf16": 1, "bf16x3": 3, "f16x2": 2}
SPLIT_KIND = {"bf16x3": 0, "f16x2": 1}


def choose_bn(cout: int) -> Tuple[int, int]:
    """N tile (multiple of 32, <= 256) and padded cout."""
    cap = int(os.environ.get("MQ_BN_CAP", "256"))          # experiment knob: narrower N tiles
    if cap < 256 and cout > cap and cout % cap == 0:
        return cap, cout
    if cout <= 256:
        bn = (cout + 31) // 32 * 32
        return bn, bn
    best = None
    for bn in (256, 224, 192, 160, 128):
        pad = (cout + bn - 1) // bn * bn
        key = (pad, -bn)
        if best is None or key < best[0]:
            best = (key, bn, pad)
    return best[1], best[2]


def pack_conv(weight: torch.Tensor, bias: Optional[torch.Tensor], kind: str, split=False,
              in_seg_stride: Optional[int] = None, on_device: bool = False,
              taps: Optional[Tuple[Sequence[int], Sequence[int]]] = None) -> PackedConv:
    """Pack a folded conv/linear weight for mq_conv_gemm.

    kind: "linear" (Cout, Cin) | "same1d" / "causal1d" (Cout, Cin, k) | "conv2d3" (Cout, Cin, 3, 3).
    split: False / "bf16" = one bf16 pass; True / "bf16x3" = the six product segments of 3-term bf16
    splits (activation expected as [x0 | x1 | x2] along channels); "f16x2" = the three product
    segments h0*g1, h1*g0, h0*g0 of 2-term fp16 splits (activation [h0 | h1]; weights pre-scaled by a
    power of two so the low term stays in fp16's normal range, undone by ``acc_scale``).  The term
    stride of the activation is ``in_seg_stride`` (default Cin).  K order is (segment, tap, channel
    chunk), matching the kernel.  ``on_device``: pack where the weight lives (the training step re-packs
    the folded weights every iteration); otherwise on the host, once at load.  "anticausal1d" (taps at
    rows 0 .. k-1) is the data gradient of a causal convolution.  "taps2d": weight (Cout, Cin, ntaps) with explicit
    per-tap (dh, dw) offsets in ``taps`` (at most MQ_MAX_TAPS; the lowered strided convolutions of the discriminators).
    """
    mode = split_mode(split)
    w = weight.detach().float()
    if not on_device:
        w = w.cpu()
    if kind == "linear":
        cout, cin = w.shape
        wt = w.reshape(cout, cin, 1)
        dh, dw = [0], [0]
    elif kind in ("same1d", "causal1d", "anticausal1d"):
        cout, cin, k = w.shape
        wt = w
        if kind == "same1d":
            if k % 2 != 1:
                raise ValueError("same-padded conv1d needs an odd kernel")
            dh = [j - (k - 1) // 2 for j in range(k)]
        elif kind == "anticausal1d":
            dh = list(range(k))
        else:
            dh = [j - (k - 1) for j in range(k)]
        dw = [0] * k
    elif kind == "conv2d3":
        cout, cin, kh, kw = w.shape
        if (kh, kw) != (3, 3):
            raise ValueError("conv2d3 expects a 3x3 kernel")
        wt = w.reshape(cout, cin, 9)
        dh = [i - 1 for i in range(3) for _ in range(3)]
        dw = [j - 1 for _ in range(3) for j in range(3)]
    elif kind == "taps2d":
        if taps is None or w.dim() != 3 or len(taps[0]) != w.shape[2] or len(taps[1]) != w.shape[2]:
            raise ValueError("taps2d expects a (Cout, Cin, ntaps) weight and taps = (dh list, dw list) of that length")
        cout, cin, _ = w.shape
        wt = w
        dh, dw = [int(v) for v in taps[0]], [int(v) for v in taps[1]]
    else:
        raise ValueError(kind)
    taps = wt.shape[2]
    if taps > _lib.MQ_MAX_TAPS:
        raise ValueError(f"{taps} taps > MQ_MAX_TAPS")
    kchunks = (cin + BLOCK_K - 1) // BLOCK_K
    cpad = kchunks * BLOCK_K
    bn, cout_pad = choose_bn(cout)
    acc_scale = 1.0
    stride = cin if in_seg_stride is None else in_seg_stride
    if mode == "bf16x3":
        w0, w1, w2 = split3_bf16(wt)
        seg_w = [w0, w1, w2, w0, w1, w0]          # smallest products first
        seg_a = [2, 1, 0, 1, 0, 0]
        a_coff = [t * stride for t in seg_a]
    elif mode == "f16x2":
        # scale so max|w| lands in [2^13, 2^14): the low term g1 (<= 2^-11 |w|) is then a normal fp16
        # number for every weight above 2^-16 of the largest, and accumulators stay far from fp32 limits
        wmax = float(wt.abs().max())
        e = 0 if wmax == 0.0 else 13 - math.floor(math.log2(wmax))
        e = max(-24, min(24, e))
        g0, g1 = split2_f16(wt * (2.0 ** e))
        acc_scale = 2.0 ** (-e)
        seg_w = [g1, g0, g0]                      # smallest products first
        seg_a = [0, 1, 0]
        a_coff = [t * stride for t in seg_a]
    else:
        seg_w = [wt.to(torch.bfloat16)]
        a_coff = [0]
    nseg = len(seg_w)
    wp = torch.zeros(cout_pad, nseg, taps, cpad, dtype=torch.float16 if mode == "f16x2" else torch.bfloat16,
                     device=w.device)
    for s, ws in enumerate(seg_w):
        wp[:cout, s, :, :cin] = ws.permute(0, 2, 1)      # (cout, taps, cin)
    wp = wp.reshape(cout_pad, nseg * taps * cpad).contiguous()
    b = None if bias is None else bias.detach().float().to(w.device).contiguous()
    return PackedConv(wp, b, cin, cout, cout_pad, bn, taps, nseg, kchunks, dh, dw,
                      a_coff + [0] * (_lib.MQ_MAX_SEGS - len(a_coff)), mode != "bf16", mode=mode,
                      acc_scale=acc_scale)


def pack_upconv(weight: torch.Tensor, bias: Optional[torch.Tensor], cx: int, cs: int) -> PackedConv:
    """Pack UpBlock.conv.conv1 (Cout, cx+cs, 3, 3) for the fused nearest-upsample + concat mode of
    mq_conv_gemm.  Output row 2i+p of conv3x3(cat[up(x), skip]) reads up(x) rows 2i+p-1 .. 2i+p+1,
    i.e. x rows {i-1, i, i} (p = 0) or {i, i, i+1} (p = 1): two row taps with pre-summed weights
    per parity (6 taps instead of 9 on the up-sampled half), plus the 9 ordinary taps on the skip."""
    w = weight.detach().float().cpu()
    cout = w.shape[0]
    if w.shape[1] != cx + cs or tuple(w.shape[2:]) != (3, 3):
        raise ValueError("pack_upconv: weight must be (Cout, cx+cs, 3, 3)")
    wx, ws = w[:, :cx], w[:, cx:]
    k1, k2 = (cx + BLOCK_K - 1) // BLOCK_K, (cs + BLOCK_K - 1) // BLOCK_K
    bn, cout_pad = choose_bn(cout)
    K = (6 * k1 + 9 * k2) * BLOCK_K
    wp = torch.zeros(2, cout_pad, K, dtype=torch.bfloat16)
    rows = {0: [wx[:, :, 0, :], wx[:, :, 1, :] + wx[:, :, 2, :]],       # half-row offsets -1, 0
            1: [wx[:, :, 0, :] + wx[:, :, 1, :], wx[:, :, 2, :]]}       # half-row offsets 0, +1
    for p in (0, 1):
        col = 0
        for rw in rows[p]:
            for j in range(3):
                wp[p, :cout, col:col + cx] = rw[:, :, j].to(torch.bfloat16)
                col += k1 * BLOCK_K
        for i in range(3):
            for j in range(3):
                wp[p, :cout, col:col + cs] = ws[:, :, i, j].to(torch.bfloat16)
                col += k2 * BLOCK_K
    dh = [-1, -1, -1, 0, 0, 0] + [i - 1 for i in range(3) for _ in range(3)]
    dh_odd = [0, 0, 0, 1, 1, 1] + [0] * 9
    dw = [-1, 0, 1, -1, 0, 1] + [j - 1 for _ in range(3) for j in range(3)]
    b = None if bias is None else bias.detach().float().cpu().contiguous()
    return PackedConv(wp.reshape(2 * cout_pad, K).contiguous(), b, cx, cout, cout_pad, bn, 15, 1, k1, dh, dw,
                      [0] * _lib.MQ_MAX_SEGS, False, up_taps=6, kchunks2=k2, cin2=cs, tap_dh_odd=dh_odd)


def choose_tile(H: int, W: int) -> Tuple[int, int]:
    """(bh, bw) with bh*bw <= 128 minimising out-of-bounds waste; prefers wide rows."""
    if W == 1:
        return 128, 1
    best = None
    for bw in range(1, min(W, 128) + 1):
        bh = min(128 // bw, 256)
        if bh < 1:
            continue
        covered = math.ceil(W / bw) * bw * math.ceil(H / bh) * bh
        util = (H * W) / covered * (bh * bw / 128.0)
        key = (-util, -bw)
        if best is None or key < best[0]:
            best = (key, bh, bw)
    return best[1], best[2]


def choose_msub(bn: int, N: int, H: int, W: int, bh: int, bw: int, up: bool = False, kblocks: int = 0,
                heavy_epilogue: bool = False) -> int:
    """Sub-tiles per CTA tile.  Narrow layers (bn <= 128) are bound by L2->SM operand traffic:
    stacking 2-4 pixel sub-tiles on one weight tile amortises the weight loads.
    Only when there are still >= 2 waves of tiles for 148 SMs."""
    if MSUB_OVERRIDE:
        m = MSUB_OVERRIDE
        return m if m * bn <= 512 else (2 if 2 * bn <= 512 else 1)
    # measured on B200 (tools/conv_bench.py): bn <= 64 -> 4; bn <= 128 -> 2 (4 for the fused up-conv);
    # bn = 256 -> 2 with a single TMEM accumulator buffer (halves the weight traffic per pixel,
    # worth more than overlapping the epilogue at K >= 2304)
    # only for long K loops (>= 64 k-blocks) whose epilogue has no residual read
    wide = 2 if (kblocks >= 64 and not heavy_epilogue) else 1
    m = 4 if bn <= 64 else ((4 if up else 2) if bn <= 128 else wide)
    while m * bn > 512:
        m //= 2
    tiles_w = math.ceil(W / bw)
    while m > 1 and N * math.ceil(H / (bh * m)) * tiles_w < 2 * 148:
        m //= 2
    return m


def choose_msub_pair(bn: int, N: int, H: int, W: int, up: bool) -> int:
    """Sub-tiles per CTA of a CTA pair (a pair tile is 2*msub sub-tiles of 16 rows x 8 columns).
    msub*bn <= 256 keeps two TMEM accumulator buffers so the epilogue overlaps the next main loop."""
    m = MSUB_PAIR_OVERRIDE if MSUB_PAIR_OVERRIDE else (4 if bn <= 64 else (2 if bn <= 128 else MSUB_PAIR_WIDE))
    if up:
        m = min(m, 2)             # the two skip-parity boxes of msub = 4 do not fit shared memory twice
    while m > 1 and m * bn > 512:
        m //= 2
    tiles_w = math.ceil(W / 8)
    while m > 1 and (H < 32 * m or N * math.ceil(H / (32 * m)) * tiles_w < 2 * 74):
        m //= 2
    return m


def conv_gemm(x: torch.Tensor, pc: PackedConv, N: int, H: int, W: int, *,
              row_mask: Optional[torch.Tensor] = None, mask_pre=False, mask_post=False,
              act=False, beta=1.0, gamma=0.5, fast_tanh=True,
              res: Optional[torch.Tensor] = None, res_mode=0, res_coff=0,
              out_f32: Optional[torch.Tensor] = None, f32_coff=0,
              out_bf16: Optional[torch.Tensor] = None, bf16_coff=0,
              out_split: Optional[torch.Tensor] = None, tile: Optional[Tuple[int, int]] = None,
              msub: Optional[int] = None, tag: str = "", x2: Optional[torch.Tensor] = None,
              halo: Optional[bool] = None, pair: Optional[bool] = None,
              out_pool: Optional[torch.Tensor] = None) -> None:
    """Launch mq_conv_gemm.  x: bf16 (fp16 for an "f16x2" weight) (N*H*W, in_ld) channel-last (any
    leading shape).  x2: skip tensor (N, 2H, W, C2) for a ``pack_upconv`` weight; outputs / masks then
    have 2H rows.  out_split: bf16 (.., 3C) or fp16 (.., 2C) multi-term output for the next split GEMM."""
    op_dt = torch.float16 if pc.mode == "f16x2" else torch.bfloat16
    _chk(x, op_dt, "x")
    in_ld = x.shape[-1]
    if x.numel() != N * H * W * in_ld:
        raise ValueError(f"x has {x.numel()} elements, expected N*H*W*in_ld = {N * H * W * in_ld}")
    p = ConvParams()
    p.inp = x.data_ptr()
    p.N, p.H, p.W, p.in_ld = N, H, W, in_ld
    p.wpack = _chk(pc.wpack, op_dt, "wpack").data_ptr()
    p.op_f16 = int(pc.mode == "f16x2")
    p.acc_scale = float(pc.acc_scale)
    p.cout, p.cout_pad, p.bn = pc.cout, pc.cout_pad, pc.bn
    p.taps, p.nseg, p.kchunks = pc.taps, pc.nseg, pc.kchunks
    for i in range(pc.taps):
        p.tap_dh[i] = pc.tap_dh[i]
        p.tap_dw[i] = pc.tap_dw[i]
    for i in range(_lib.MQ_MAX_SEGS):
        p.a_coff[i] = pc.a_coff[i]
    hm = 1
    if pc.up_taps:
        if x2 is None:
            raise ValueError("this packed weight needs the skip tensor x2")
        _chk(x2, torch.bfloat16, "x2")
        if x2.numel() != N * 2 * H * W * x2.shape[-1]:
            raise ValueError("x2 must be (N, 2H, W, C2)")
        p.in2, p.in2_ld, p.up_taps, p.kchunks2 = x2.data_ptr(), x2.shape[-1], pc.up_taps, pc.kchunks2
        for i in range(pc.taps):
            p.tap_dh_odd[i] = pc.tap_dh_odd[i]
        hm = 2
    conv3 = (pc.taps == 9 and not pc.up_taps
             and list(pc.tap_dh[:9]) == [-1, -1, -1, 0, 0, 0, 1, 1, 1] and list(pc.tap_dw[:9]) == [-1, 0, 1] * 3)
    conv1d = (W == 1 and not pc.up_taps and all(d == 0 for d in pc.tap_dw[:pc.taps])
              and all(pc.tap_dh[i] == pc.tap_dh[0] + i for i in range(pc.taps)))
    if pair is None and conv1d:
        # row-halo CTA-pair loop for the wide 1-D layers (encoder / decoder blocks): measured against the
        # tap-shifted loop in tools/conv_bench.py
        pair = (PAIR_DEFAULT and PAIR_1D and pc.bn >= PAIR_1D_MIN_BN and pc.bn % 32 == 0 and H >= 256
                and tile is None and halo is not True and (msub is None or msub == 1))
    if pair and conv1d:
        if halo:
            raise ValueError("pair and halo main loops are exclusive")
        p.pair, p.halo = 1, 0
        halo = False
        tile = (128, 1)
        msub = 1
    if pair is None:
        # measured (tools/conv_bench.py, profiles/conv_bench_r01_*.log): the CTA-pair loop beats the
        # single-CTA halo / tap loops on every refiner layer shape, plain and fused up-conv
        pair = (PAIR_DEFAULT and (conv3 or bool(pc.up_taps)) and pc.bn > PAIR_MIN_BN and W >= 8 and H >= 32
                and tile is None and halo is not True and pc.bn % 32 == 0)
    if halo is None:
        halo = HALO_DEFAULT and conv3 and pc.nseg == 1 and W >= 8 and tile is None and not pair
    if pair and halo:
        raise ValueError("pair and halo main loops are exclusive")
    p.pair = int(bool(pair))
    if pair and conv1d:
        bh, bw = tile
    elif pair:
        bh, bw = 16, 8
        if msub is None:
            msub = choose_msub_pair(pc.bn, N, H, W, bool(pc.up_taps))
    elif halo:
        bh, bw = 16, 8
    else:
        bh, bw = tile if tile is not None else choose_tile(H, W)
    p.halo = int(bool(halo))
    p.bh, p.bw = bh, bw
    if msub is None:
        kblocks = pc.wpack.shape[1] // BLOCK_K
        msub = choose_msub(pc.bn, N, H, W, bh, bw, bool(pc.up_taps), kblocks,
                           heavy_epilogue=(res_mode != 0 or out_split is not None or out_f32 is not None))
    p.msub = msub
    p.bias = _ptr(pc.bias)
    if row_mask is not None:
        _chk(row_mask, torch.uint8, "row_mask")
        if row_mask.numel() != N * H * hm:
            raise ValueError("row_mask must have one entry per output row")
    p.row_mask = _ptr(row_mask)
    p.mask_pre, p.mask_post = int(mask_pre), int(mask_post)
    p.act, p.fast_tanh = int(act), int(fast_tanh)
    p.beta, p.gamma = float(beta), float(gamma)
    p.res_mode = int(res_mode)
    if res is not None:
        if res.dtype not in (torch.bfloat16, torch.float32):
            raise TypeError("res must be bf16 or fp32")
        p.res = res.data_ptr()
        p.res_is_bf16 = int(res.dtype == torch.bfloat16)
        p.res_ld = res.shape[-1]
        p.res_coff = res_coff
    if out_f32 is not None:
        _chk(out_f32, torch.float32, "out_f32")
        p.out_f32, p.f32_ld, p.f32_coff = out_f32.data_ptr(), out_f32.shape[-1], f32_coff
    if out_bf16 is not None:
        _chk(out_bf16, torch.bfloat16, "out_bf16")
        p.out_bf16, p.bf16_ld, p.bf16_coff = out_bf16.data_ptr(), out_bf16.shape[-1], bf16_coff
    if out_pool is not None:
        _chk(out_pool, torch.bfloat16, "out_pool")
        if H % 2 or out_pool.numel() != N * (H // 2) * W * out_pool.shape[-1]:
            raise ValueError("out_pool must be (N, H/2, W, C) with H even")
        p.out_pool, p.pool_ld = out_pool.data_ptr(), out_pool.shape[-1]
    if out_split is not None:
        nt = _split_terms_of(out_split)
        if out_split.shape[-1] % nt:
            raise ValueError(f"out_split last dim must be {nt}*C")
        p.out_split, p.split_ld, p.split_seg = out_split.data_ptr(), out_split.shape[-1], out_split.shape[-1] // nt
        p.split_kind = int(nt == 2)
    meta = None
    if _lib.profiler is not None or _lib.NVTX:
        pix = float(N) * H * W * hm
        if pc.up_taps:
            meta = {"tag": tag, "flops": 2.0 * pix * pc.cout * (pc.cin + pc.cin2) * 9,  # the reference's conv
                    "mma_flops": 2.0 * pix * pc.cout_pad * pc.wpack.shape[1]}
        else:
            meta = {"tag": tag, "flops": 2.0 * pix * pc.cout * pc.cin * pc.taps,        # algorithmic
                    "mma_flops": 2.0 * pix * pc.cout_pad * pc.taps * pc.nseg * pc.kchunks * BLOCK_K}  # issued
    _lib.call("mq_conv_gemm", C.byref(p), _stream(), meta=meta)


def _split_terms_of(out_split: torch.Tensor) -> int:
    """A multi-term output buffer is bf16 (three terms) or fp16 (two terms)."""
    if out_split.dtype == torch.bfloat16:
        nt = 3
    elif out_split.dtype == torch.float16:
        nt = 2
    else:
        raise TypeError(f"out_split: expected bf16 (bf16x3) or fp16 (f16x2), got {out_split.dtype}")
    if not out_split.is_cuda or not out_split.is_contiguous():
        raise ValueError("out_split: expected a contiguous CUDA tensor")
    return nt


# ---------------------------------------------------------------------------
# training step: data and weight gradients of the convolutions
# ---------------------------------------------------------------------------
DGRAD_KIND = {"linear": "linear", "same1d": "same1d", "causal1d": "anticausal1d", "conv2d3": "conv2d3"}


def dgrad_weight(weight: torch.Tensor, kind: str, taps=None):
    """The convolution whose forward pass is the data gradient of ``conv(x, weight)``: taps mirrored,
    in/out channels swapped.  dx = conv(dy, w'), w'[ci, co, j'] = w[co, ci, k-1-j']; a causal
    convolution (taps at rows -(k-1) .. 0, attentions.py:471-474) turns anti-causal (rows 0 .. k-1)."""
    if kind == "linear":
        return weight.t().contiguous(), "linear"
    if kind in ("same1d", "causal1d"):
        return weight.flip(2).transpose(0, 1).contiguous(), DGRAD_KIND[kind]
    if kind == "conv2d3":
        return weight.flip(2, 3).transpose(0, 1).contiguous(), "conv2d3"
    if kind == "taps2d":                       # same tap order, every offset negated: returns (weight, kind, taps)
        return weight.transpose(0, 1).contiguous(), "taps2d", ([-int(v) for v in taps[0]], [-int(v) for v in taps[1]])
    raise ValueError(kind)


def conv_taps(kind: str, wshape, taps=None) -> Tuple[List[int], List[int]]:
    """(tap_dh, tap_dw) of a convolution kind in the weight's own tap order (as pack_conv)."""
    if kind == "taps2d":
        return [int(v) for v in taps[0]], [int(v) for v in taps[1]]
    if kind == "linear":
        return [0], [0]
    if kind == "same1d":
        k = wshape[2]
        return [j - (k - 1) // 2 for j in range(k)], [0] * k
    if kind == "causal1d":
        k = wshape[2]
        return [j - (k - 1) for j in range(k)], [0] * k
    if kind == "conv2d3":
        return [i - 1 for i in range(3) for _ in range(3)], [j - 1 for _ in range(3) for j in range(3)]
    raise ValueError(kind)


def conv_wgrad(dy: torch.Tensor, x: torch.Tensor, N: int, H: int, W: int, cout: int, cin: int,
               tap_dh: Sequence[int], tap_dw: Sequence[int], *, split: Optional[int] = None,
               tag: str = "") -> torch.Tensor:
    """Launch mq_conv_wgrad: dy bf16 (N,H,W,>=cout), x bf16 (N,H,W,>=cin) channel-last ->
    fp32 (taps, cout, cin) = sum over pixels of dy[p, co] * x[p + tap, ci]."""
    _chk(dy, torch.bfloat16, "dy")
    _chk(x, torch.bfloat16, "x")
    p = _lib.WgradParams()
    p.dy, p.dy_ld = dy.data_ptr(), dy.shape[-1]
    p.x, p.x_ld = x.data_ptr(), x.shape[-1]
    if dy.numel() != N * H * W * p.dy_ld or x.numel() != N * H * W * p.x_ld:
        raise ValueError("dy / x must hold N*H*W pixels")
    p.N, p.H, p.W = N, H, W
    p.cout, p.cin = cout, cin
    taps = len(tap_dh)
    p.taps = taps
    for i in range(taps):
        p.tap_dh[i] = tap_dh[i]
        p.tap_dw[i] = tap_dw[i]
    p.bh, p.bw = (64, 1) if W == 1 else (8, 8)
    p.split = 1
    if split is None:
        split = _lib.lib().mq_conv_wgrad_split(C.byref(p))
    p.split = max(1, int(split))
    part = torch.empty(p.split, taps, cout, cin, dtype=torch.float32, device=dy.device)
    p.dw = part.data_ptr()
    meta = None
    if _lib.profiler is not None:
        fl = 2.0 * N * H * W * cout * cin * taps
        meta = {"tag": tag, "flops": fl, "mma_flops": fl}
    _lib.call("mq_conv_wgrad", C.byref(p), _stream(), meta=meta)
    return part[0] if p.split == 1 else part.sum(dim=0)


def act_forward(u: torch.Tensor, res: Optional[torch.Tensor], row_mask: Optional[torch.Tensor], pix_per_row: int,
                beta: float = 1.0, gamma: float = 0.5) -> torch.Tensor:
    """u (.., C) fp32 conv output -> bf16 aptx(u) [+ res], zero at padded rows: mq_act_forward."""
    _chk(u, torch.float32, "u")
    Cc = u.shape[-1]
    pixels = u.numel() // Cc
    if res is not None:
        _chk(res, torch.bfloat16, "res")
        if res.numel() != u.numel():
            raise ValueError("res must have the shape of u")
    if row_mask is not None:
        _chk(row_mask, torch.uint8, "row_mask")
        if row_mask.numel() * pix_per_row != pixels:
            raise ValueError("row_mask must have one entry per row of pix_per_row pixels")
    out = torch.empty(u.shape, dtype=torch.bfloat16, device=u.device)
    _lib.call("mq_act_forward", u.data_ptr(), _ptr(res), _ptr(row_mask), pixels, Cc, pix_per_row, float(beta), float(gamma),
              out.data_ptr(), _stream())
    return out


def act_backward(dy: torch.Tensor, u: torch.Tensor, row_mask: Optional[torch.Tensor], pix_per_row: int,
                 beta: float = 1.0, gamma: float = 0.5, want_res: bool = False, want_bias: bool = False):
    """-> (du, dres or None[, dbias fp32 (C)]), du / dres bf16: mq_act_backward.  ``want_bias``: also the column sums of
    du (the producing convolution's bias gradient), reduced in the same pass instead of re-reading du."""
    _chk(dy, torch.bfloat16, "dy")
    _chk(u, torch.float32, "u")
    if dy.numel() != u.numel():
        raise ValueError("dy must have the shape of u")
    Cc = u.shape[-1]
    pixels = u.numel() // Cc
    du = torch.empty(u.shape, dtype=torch.bfloat16, device=u.device)
    dres = torch.empty_like(du) if want_res else None
    part = None
    if want_bias:
        nb = _lib.lib().mq_act_bias_blocks(pixels, Cc)
        if nb > 0:
            part = torch.empty(nb, Cc, dtype=torch.float32, device=u.device)
    _lib.call("mq_act_backward", dy.data_ptr(), u.data_ptr(), _ptr(row_mask), pixels, Cc, pix_per_row, float(beta),
              float(gamma), du.data_ptr(), _ptr(dres), _ptr(part), _stream())
    if want_bias:
        db = part.sum(dim=0) if part is not None else du.reshape(-1, Cc).sum(dim=0, dtype=torch.float32)
        return du, dres, db
    return du, dres


def _chk_cl(t: torch.Tensor, name: str) -> torch.Tensor:
    """A (B, C, H, W) tensor stored channels-last (the discriminators' cuDNN layout)."""
    if not t.is_cuda or t.dim() != 4 or not t.is_contiguous(memory_format=torch.channels_last):
        raise ValueError(f"{name}: expected a channels_last (B, C, H, W) CUDA tensor")
    if t.shape[1] % 8:
        raise ValueError(f"{name}: channel count must be a multiple of 8")
    return t


def leaky_mask_forward(y: torch.Tensor, pix_mask: torch.Tensor, slope: float = 0.2,
                       bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y (B, C, H, W) channels_last fp32 / bf16; pix_mask (B, H, W) uint8 (1 = padded); bias fp32 (C) or None -> bf16
    LeakyReLU(y + bias) with padded pixels zeroed (discriminators.py:234, 247): mq_leaky_mask_forward."""
    _chk_cl(y, "y")
    if y.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("y must be fp32 or bf16")
    _chk(pix_mask, torch.uint8, "pix_mask")
    B, Cc, H, W = y.shape
    if pix_mask.numel() != B * H * W:
        raise ValueError("pix_mask must have one entry per pixel")
    if bias is not None:
        _chk(bias, torch.float32, "bias")
    out = torch.empty_like(y, dtype=torch.bfloat16)
    _lib.call("mq_leaky_mask_forward", y.data_ptr(), int(y.dtype == torch.bfloat16), _ptr(bias), pix_mask.data_ptr(),
              B * H * W, Cc, float(slope), out.data_ptr(), _stream())
    return out


def leaky_mask_backward(dout: torch.Tensor, y: torch.Tensor, pix_mask: torch.Tensor, slope: float = 0.2,
                        bias: Optional[torch.Tensor] = None, want_bias: bool = False):
    """Gradient of leaky_mask_forward w.r.t. y (bf16, y's layout) [and w.r.t. bias, fp32 (C)]: mq_leaky_mask_backward."""
    _chk_cl(y, "y")
    _chk_cl(dout, "dout")
    if dout.dtype != torch.bfloat16:
        raise TypeError("dout must be bf16")
    B, Cc, H, W = y.shape
    du = torch.empty_like(y, dtype=torch.bfloat16)
    part = None
    if want_bias:
        nb = _lib.lib().mq_act_bias_blocks(B * H * W, Cc)
        if nb > 0:
            part = torch.empty(nb, Cc, dtype=torch.float32, device=y.device)
    _lib.call("mq_leaky_mask_backward", dout.data_ptr(), y.data_ptr(), int(y.dtype == torch.bfloat16), _ptr(bias),
              pix_mask.data_ptr(), B * H * W, Cc, float(slope), du.data_ptr(), _ptr(part), _stream())
    if want_bias:
        db = part.sum(dim=0) if part is not None else du.float().sum(dim=(0, 2, 3))
        return du, db
    return du


def cb2d_point_forward(s: torch.Tensor, wpw, bpw, wout, bout, row_mask: Optional[torch.Tensor],
                       fast_tanh: bool = False) -> torch.Tensor:
    """s (B, T, C) fp32 (masked depth-wise output) -> y (B, T, C): mq_cb2d_point_forward."""
    _chk(s, torch.float32, "s")
    Cc = s.shape[-1]
    rows = s.numel() // Cc
    prm = [_chk(t.detach().contiguous(), torch.float32, n) for t, n in ((wpw, "wpw"), (bpw, "bpw"), (wout, "wout"), (bout, "bout"))]
    if row_mask is not None:
        _chk(row_mask, torch.uint8, "row_mask")
    y = torch.empty_like(s)
    _lib.call("mq_cb2d_point_forward", s.data_ptr(), _ptr(row_mask), rows, Cc, prm[0].data_ptr(), prm[1].data_ptr(),
              prm[2].data_ptr(), prm[3].data_ptr(), int(fast_tanh), y.data_ptr(), _stream())
    return y


def cb2d_point_backward(s: torch.Tensor, dy: torch.Tensor, wpw, bpw, wout, row_mask: Optional[torch.Tensor],
                        fast_tanh: bool = False):
    """-> (ds, dwpw, dbpw, dwout, dbout) for y = cb2d_point_forward(s, ...): mq_cb2d_backward (two launches)."""
    _chk(s, torch.float32, "s")
    _chk(dy, torch.float32, "dy")
    Cc = s.shape[-1]
    rows = s.numel() // Cc
    prm = [_chk(t.detach().contiguous(), torch.float32, n) for t, n in ((wpw, "wpw"), (bpw, "bpw"), (wout, "wout"))]
    nb = _lib.lib().mq_cb2d_grad_blocks(rows, Cc)
    ds = torch.empty_like(s)
    part = torch.empty(nb, 3, Cc, dtype=torch.float32, device=s.device)
    _lib.call("mq_cb2d_backward", s.data_ptr(), dy.data_ptr(), _ptr(row_mask), rows, Cc, prm[0].data_ptr(),
              prm[1].data_ptr(), prm[2].data_ptr(), int(fast_tanh), ds.data_ptr(), part.data_ptr(), _stream())
    g = part.sum(dim=0)
    return ds, g[0], g[1], g[2], dy.sum().reshape(1)


# ---------------------------------------------------------------------------
# element-wise / reduction entry points
# ---------------------------------------------------------------------------
def split_bf16(x: torch.Tensor, nterms: int) -> torch.Tensor:
    """fp32 -> 16-bit operand terms along channels: 1 = bf16, 3 = bf16x3, 2 = f16x2 (fp16 tensor)."""
    _chk(x, torch.float32, "x")
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    out = torch.empty(*x.shape[:-1], nterms * Cc, dtype=torch.float16 if nterms == 2 else torch.bfloat16,
                      device=x.device)
    _lib.call("mq_split_bf16", x.data_ptr(), out.data_ptr(), rows, Cc, nterms, _stream())
    return out


def convblock2d(x: torch.Tensor, B: int, T: int, Cc: int, dw: torch.Tensor, pw: torch.Tensor, bout: float,
                row_mask: Optional[torch.Tensor], fast_tanh: bool, *, out_f32=None, out_bf16=None,
                out_split=None, table: Optional[torch.Tensor] = None, table_off: int = 0,
                table_inv_h: float = 0.0) -> None:
    p = Cb2dParams()
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("x must be fp32 or bf16")
    p.x, p.x_is_bf16 = x.data_ptr(), int(x.dtype == torch.bfloat16)
    p.B, p.T, p.C = B, T, Cc
    p.dw, p.pw = _chk(dw, torch.float32, "dw").data_ptr(), _chk(pw, torch.float32, "pw").data_ptr()
    p.bout = float(bout)
    p.row_mask = _ptr(row_mask)
    p.fast_tanh = int(fast_tanh)
    p.out_f32, p.out_bf16, p.out_split = _ptr(out_f32), _ptr(out_bf16), _ptr(out_split)
    if out_split is not None:
        p.split_kind = int(_split_terms_of(out_split) == 2)
    if table is not None:
        _chk(table, torch.float32, "table")
        p.table, p.table_n, p.table_off, p.table_inv_h = table.data_ptr(), table.shape[0], int(table_off), float(table_inv_h)
    _lib.call("mq_convblock2d", C.byref(p), _stream())


def cam_chunks(T: int) -> int:
    return _lib.lib().mq_cam_chunks(T)


def cam_gate(o: torch.Tensor, row_mask: Optional[torch.Tensor], B: int, T: int, Cc: int,
             w0, b0, w2, b2) -> torch.Tensor:
    _chk(o, torch.float32, "o")
    nchunk = cam_chunks(T)
    part = torch.empty(B, nchunk, 2, Cc, dtype=torch.float32, device=o.device)
    _lib.call("mq_cam_reduce", o.data_ptr(), _ptr(row_mask), B, T, Cc, part.data_ptr(), _stream())
    gate = torch.empty(B, Cc, dtype=torch.float32, device=o.device)
    _lib.call("mq_cam_gate", part.data_ptr(), _ptr(row_mask), B, T, Cc, w0.shape[0], w0.data_ptr(),
              b0.data_ptr(), w2.data_ptr(), b2.data_ptr(), gate.data_ptr(), _stream())
    return gate


def cbam_apply(o, gate, res, row_mask, B, T, Cc, sam_w, beta, gamma, *, out_f32=None, out_bf16=None,
               out_split=None) -> None:
    p = CbamApplyParams()
    p.o, p.gate, p.res = _chk(o, torch.float32, "o").data_ptr(), gate.data_ptr(), _chk(res, torch.float32, "res").data_ptr()
    p.row_mask = _ptr(row_mask)
    p.B, p.T, p.C = B, T, Cc
    p.sam_w = _chk(sam_w, torch.float32, "sam_w").data_ptr()
    p.beta, p.gamma = float(beta), float(gamma)
    p.out_f32, p.out_bf16, p.out_split = _ptr(out_f32), _ptr(out_bf16), _ptr(out_split)
    if out_split is not None:
        p.split_kind = int(_split_terms_of(out_split) == 2)
    _lib.call("mq_cbam_apply", C.byref(p), _stream())


def fsq_params(levels: Sequence[int]) -> FsqParams:
    """FSQ constants exactly as quantizer.py:68-72,109-114,132 computes them (fp32)."""
    f = FsqParams()
    D = len(levels)
    if D > 8:
        raise ValueError("at most 8 FSQ dims")
    lv = torch.tensor(list(levels), dtype=torch.int32)
    half_l = (lv - 1) * (1 + 1e-3) / 2
    offset = torch.where(lv % 2 == 0, 0.5, 0.0)
    shift = (offset / half_l).atanh()
    basis = torch.cumprod(torch.tensor([1] + list(levels[:-1])), dim=0)
    f.D = D
    for i in range(D):
        f.half_l[i] = float(half_l[i])
        f.offset[i] = float(offset[i])
        f.shift[i] = float(shift[i])
        f.half_w[i] = int(levels[i]) // 2
        f.basis[i] = int(basis[i])
        f.levels[i] = int(levels[i])
    return f


def qin_fsq(y: torch.Tensor, w: torch.Tensor, b: torch.Tensor, fsq: FsqParams, want_z=False):
    _chk(y, torch.float32, "y")
    Cc = y.shape[-1]
    rows = y.numel() // Cc
    idx = torch.empty(y.shape[:-1], dtype=torch.int64, device=y.device)
    z = torch.empty(*y.shape[:-1], fsq.D, dtype=torch.float32, device=y.device) if want_z else None
    _lib.call("mq_qin_fsq", y.data_ptr(), rows, Cc, _chk(w, torch.float32, "w").data_ptr(),
              _chk(b, torch.float32, "b").data_ptr(), C.byref(fsq), idx.data_ptr(), _ptr(z), _stream())
    return (idx, z) if want_z else idx


def fsq_quantize(z: torch.Tensor, fsq: FsqParams, want_codes=False):
    _chk(z, torch.float32, "z")
    rows = z.numel() // fsq.D
    idx = torch.empty(z.shape[:-1], dtype=torch.int64, device=z.device)
    codes = torch.empty_like(z) if want_codes else None
    _lib.call("mq_fsq_quantize", z.data_ptr(), rows, C.byref(fsq), idx.data_ptr(), _ptr(codes), _stream())
    return (idx, codes) if want_codes else idx


@dataclass
class PackedCodebook:
    """Codebook operand of mq_vq_nearest, packed once (host side)."""
    cb_img: torch.Tensor       # uint8 shared-memory images [tiles][nterm][256][128]
    c2: torch.Tensor           # fp32 [k_pad], +inf on padding codes
    codebook: torch.Tensor     # fp32 (k, d)
    k: int
    k_pad: int
    d: int
    mode: int                  # 0 = bf16, 1 = f16x2
    acc_scale: float
    fold: bool = False         # ||c||^2 and the factor -2 live inside the GEMM (three augmented K columns per code)
    zconst: float = 1.0        # what the kernel writes into those columns of the latent operand

    def to(self, device):
        self.cb_img, self.c2, self.codebook = self.cb_img.to(device), self.c2.to(device), self.codebook.to(device)
        return self


def pack_codebook(codebook: torch.Tensor, precision: str = "f16x2", fold: Optional[bool] = None) -> PackedCodebook:
    """Pack a (K, D <= 64) fp32 codebook for mq_vq_nearest: 16-bit operand terms laid out as the
    128-byte-swizzled [256 codes][64 K] shared-memory tiles the tensor core reads (slices = 4/ks codes side
    by side per 128-byte row when D <= 32), plus ||c||^2 in fp32 (computed in float64).

    ``fold`` (default: whenever D + 3 fits the code's K-steps): the score ||c||^2 - 2 z.c comes out of the GEMM itself.
    The codebook is stored as -2 c and three augmented K columns D .. D+2 hold multi-term splits of ||c||^2 / zconst
    (f16x2: two 2-term fp16 splits = 44 bits; bf16: three bf16 terms = 24 bits), multiplied in the kernel by the
    constant zconst = 2^p written into the latent operand; p keeps the fp16 terms in range."""
    if precision not in ("f16x2", "bf16"):
        raise ValueError("precision must be 'f16x2' or 'bf16'")
    cb = codebook.detach().float().cpu().contiguous()
    K, D = cb.shape
    if not 1 <= D <= 64:
        raise ValueError("codebook width must be in [1, 64]")
    ks = 1 if D <= 16 else (2 if D <= 32 else 4)
    if fold is None:
        fold = D + 3 <= 16 * ks
    if fold and D + 3 > 16 * ks:
        raise ValueError("fold needs D + 3 <= 16 * ks columns")
    slices = 4 // ks
    per_tile = 256 * slices
    k_pad = (K + per_tile - 1) // per_tile * per_tile
    tiles = k_pad // per_tile
    acc_scale, zconst = 1.0, 1.0
    c2_64 = (cb.double() ** 2).sum(1)
    src = cb * (-2.0) if fold else cb                             # exact: a power of two
    aug: List[torch.Tensor] = []                                  # per term: (K, 3) augmented columns
    if precision == "f16x2":
        cmax = float(src.abs().max())
        e = 0 if cmax == 0.0 else 13 - math.floor(math.log2(cmax))
        e = max(-24, min(24, e))
        terms = list(split2_f16(src * (2.0 ** e)))                # g0, g1
        acc_scale = 2.0 ** (-e)
        dt = torch.float16
        if fold:
            c2s = c2_64 * (2.0 ** e)                              # in accumulator units
            top = float(c2s.max())
            p = 0 if top <= 0 else max(0, math.ceil(math.log2(top)) - 14)
            if p > 15:
                raise ValueError("codebook norms too large for the folded fp16 form; pass fold=False")
            zconst = 2.0 ** p
            v = c2s / zconst
            a0, b0 = split2_f16(v.float())
            r = v - a0.double() - b0.double()
            a1, b1 = split2_f16(r.float())
            zero = torch.zeros_like(a0)
            aug = [torch.stack([a0, a1, zero], 1), torch.stack([b0, b1, zero], 1)]
    else:
        terms = [src.to(torch.bfloat16)]
        dt = torch.bfloat16
        if fold:
            t0, t1, t2 = split3_bf16(c2_64.float())
            aug = [torch.stack([t0, t1, t2], 1)]
    dense = torch.zeros(tiles, len(terms), 256, 64, dtype=dt)          # [tile][term][row n][K column]
    code = torch.arange(k_pad)
    t, rem = code // per_tile, code % per_tile
    s, n = rem // 256, rem % 256                                       # code = (tile*slices + slice)*256 + row
    valid = code < K
    for j, tj in enumerate(terms):
        for i in range(D):
            dense[t[valid], j, n[valid], (s[valid] * 16 * ks + i)] = tj[code[valid], i]
        if fold:
            for i in range(3):
                dense[t[valid], j, n[valid], (s[valid] * 16 * ks + D + i)] = aug[j][code[valid], i].to(dt)
    # 128-byte swizzle of a 1024-byte-aligned tile: 16-byte chunk c of row n lives at chunk c ^ (n & 7)
    chunks = dense.view(tiles, len(terms), 256, 8, 8)
    rows = torch.arange(256)
    srcc = (torch.arange(8)[None, :] ^ (rows[:, None] & 7))             # image chunk p holds data chunk p ^ (n&7)
    img = torch.gather(chunks, 3, srcc[None, None, :, :, None].expand(tiles, len(terms), 256, 8, 8))
    c2 = torch.full((k_pad,), float("inf"), dtype=torch.float32)
    c2[:K] = c2_64.float()
    return PackedCodebook(img.contiguous().view(torch.uint8).reshape(-1), c2, cb, K, k_pad, D,
                          1 if precision == "f16x2" else 0, acc_scale, bool(fold), float(zconst))


def vq_nearest(z: torch.Tensor, pc: PackedCodebook, want_codes: bool = True, want_dist: bool = False):
    """Nearest codeword of every row of z (n, d) fp32: returns idx (n,) int64 [, codes (n, d)] [, dist (n,)]."""
    _chk(z, torch.float32, "z")
    if z.dim() != 2 or z.shape[1] != pc.d:
        raise ValueError(f"z must be (n, {pc.d})")
    n = z.shape[0]
    idx = torch.empty(n, dtype=torch.int64, device=z.device)
    codes = torch.empty(n, pc.d, dtype=torch.float32, device=z.device) if want_codes else None
    dist = torch.empty(n, dtype=torch.float32, device=z.device) if want_dist else None
    p = VqParams()
    p.z, p.n, p.d = z.data_ptr(), n, pc.d
    p.cb_img = _chk(pc.cb_img, torch.uint8, "cb_img").data_ptr()
    p.c2 = _chk(pc.c2, torch.float32, "c2").data_ptr()
    p.codebook = _chk(pc.codebook, torch.float32, "codebook").data_ptr()
    p.k, p.k_pad, p.mode, p.acc_scale = pc.k, pc.k_pad, pc.mode, float(pc.acc_scale)
    p.idx, p.codes_out, p.dist_out = idx.data_ptr(), _ptr(codes), _ptr(dist)
    p.fold, p.zconst = int(pc.fold), float(pc.zconst)
    meta = None
    if _lib.profiler is not None:
        meta = {"tag": f"vq k={pc.k} d={pc.d}", "flops": 2.0 * n * pc.k * pc.d}
    _lib.call("mq_vq_nearest", C.byref(p), _stream(), meta=meta)
    out = [idx]
    if want_codes:
        out.append(codes)
    if want_dist:
        out.append(dist)
    return out[0] if len(out) == 1 else tuple(out)


def code_gather(idx: torch.Tensor, table: torch.Tensor, *, bf16=True, f32=False, bad: Optional[torch.Tensor] = None):
    _chk(idx, torch.int64, "idx")
    _chk(table, torch.float32, "table")
    rows, Cc = idx.numel(), table.shape[1]
    ob = torch.empty(*idx.shape, Cc, dtype=torch.bfloat16, device=idx.device) if bf16 else None
    of = torch.empty(*idx.shape, Cc, dtype=torch.float32, device=idx.device) if f32 else None
    _lib.call("mq_code_gather", idx.data_ptr(), rows, table.data_ptr(), table.shape[0], Cc, _ptr(ob), _ptr(of),
              _ptr(bad), _stream())
    return ob, of


def sequence_mask(lengths: torch.Tensor, T: int) -> torch.Tensor:
    _chk(lengths, torch.int64, "lengths")
    B = lengths.numel()
    m = torch.empty(B, T, dtype=torch.uint8, device=lengths.device)
    _lib.call("mq_sequence_mask", lengths.data_ptr(), B, T, m.data_ptr(), _stream())
    return m


def refiner_level_rows(T: int, depth: int) -> Tuple[int, List[int]]:
    mult = 1 << depth
    T8 = (T + mult - 1) // mult * mult
    return T8, [T8 >> l for l in range(depth + 1)]


def refiner_masks(mask: Optional[torch.Tensor], B: int, T: int, depth: int, device):
    """Returns (T8, down[l], up[l]) with per-level (B, H_l) uint8 views."""
    T8, rows = refiner_level_rows(T, depth)
    total = B * sum(rows)
    down = torch.empty(total, dtype=torch.uint8, device=device)
    up = torch.empty(total, dtype=torch.uint8, device=device)
    _lib.call("mq_refiner_masks", _ptr(mask), B, T, depth, down.data_ptr(), up.data_ptr(), _stream())
    dl, ul, off = [], [], 0
    for h in rows:
        dl.append(down[off:off + B * h].view(B, h))
        ul.append(up[off:off + B * h].view(B, h))
        off += B * h
    return T8, dl, ul


def zero_rows(x: torch.Tensor, mask_new: torch.Tensor, mask_old: Optional[torch.Tensor]) -> None:
    """In place: zero rows padded under mask_new but not under mask_old.  x: (rows, ...) bf16."""
    rows = mask_new.numel()
    row_bytes = x.numel() // rows * x.element_size()
    _lib.call("mq_zero_rows", x.data_ptr(), _chk(mask_new, torch.uint8, "mask_new").data_ptr(), _ptr(mask_old),
              rows, row_bytes, _stream())


def avgpool_mask(x: torch.Tensor, mask_out: Optional[torch.Tensor], B, H, F, Cc) -> torch.Tensor:
    _chk(x, torch.bfloat16, "x")
    y = torch.empty(B, H // 2, F, Cc, dtype=torch.bfloat16, device=x.device)
    _lib.call("mq_avgpool_mask", x.data_ptr(), y.data_ptr(), _ptr(mask_out), B, H, F, Cc, _stream())
    return y


def upcat_mask(x: torch.Tensor, skip: torch.Tensor, mask_out: Optional[torch.Tensor], B, H, F, Cx, Cs) -> torch.Tensor:
    _chk(x, torch.bfloat16, "x")
    _chk(skip, torch.bfloat16, "skip")
    y = torch.empty(B, H, F, Cx + Cs, dtype=torch.bfloat16, device=x.device)
    _lib.call("mq_upcat_mask", x.data_ptr(), skip.data_ptr(), y.data_ptr(), _ptr(mask_out), B, H, F, Cx, Cs, _stream())
    return y


def refiner_stem(r: torch.Tensor, mask: Optional[torch.Tensor], B, T, T8, F, Cc, w, b, fast_tanh) -> torch.Tensor:
    _chk(r, torch.float32, "r")
    y = torch.empty(B, T8, F, Cc, dtype=torch.bfloat16, device=r.device)
    _lib.call("mq_refiner_stem", r.data_ptr(), _ptr(mask), B, T, T8, F, Cc, w.data_ptr(), b.data_ptr(),
              int(fast_tanh), y.data_ptr(), _stream())
    return y


def avgpool_mask_split(x: torch.Tensor, mask_out: Optional[torch.Tensor], B, H, F, Cc) -> torch.Tensor:
    """fp32-grade decoder mode: x fp16 (B, H, F, 2C) two-term -> (B, H/2, F, 2C)."""
    _chk(x, torch.float16, "x")
    y = torch.empty(B, H // 2, F, 2 * Cc, dtype=torch.float16, device=x.device)
    _lib.call("mq_avgpool_mask_split", x.data_ptr(), y.data_ptr(), _ptr(mask_out), B, H, F, Cc, _stream())
    return y


def upcat_mask_split(x: torch.Tensor, skip: torch.Tensor, mask_out: Optional[torch.Tensor], B, H, F, Cx, Cs) -> torch.Tensor:
    """fp32-grade decoder mode: x (B, H/2, F, 2Cx), skip (B, H, F, 2Cs) fp16 -> (B, H, F, 2(Cx+Cs)) = [x0|s0|x1|s1]."""
    _chk(x, torch.float16, "x")
    _chk(skip, torch.float16, "skip")
    y = torch.empty(B, H, F, 2 * (Cx + Cs), dtype=torch.float16, device=x.device)
    _lib.call("mq_upcat_mask_split", x.data_ptr(), skip.data_ptr(), y.data_ptr(), _ptr(mask_out), B, H, F, Cx, Cs, _stream())
    return y


def refiner_stem_split(r: torch.Tensor, mask: Optional[torch.Tensor], B, T, T8, F, Cc, w, b) -> torch.Tensor:
    _chk(r, torch.float32, "r")
    y = torch.empty(B, T8, F, 2 * Cc, dtype=torch.float16, device=r.device)
    _lib.call("mq_refiner_stem_split", r.data_ptr(), _ptr(mask), B, T, T8, F, Cc, w.data_ptr(), b.data_ptr(),
              y.data_ptr(), _stream())
    return y


def refiner_tail(taps: torch.Tensor, mask: Optional[torch.Tensor], B, T, T8, F, bias: float, reproj_t, M,
                 r: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _chk(taps, torch.float32, "taps")
    if out is None:
        out = torch.empty(B, T, M, dtype=torch.float32, device=taps.device)
    _lib.call("mq_refiner_tail", taps.data_ptr(), taps.shape[-1], _ptr(mask), B, T, T8, F, float(bias),
              reproj_t.data_ptr(), M, r.data_ptr(), out.data_ptr(), _stream())
    return out
