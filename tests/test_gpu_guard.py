"""Out-of-bounds WRITE check of every kernel family without compute-sanitizer (the tool is closed on this GPU pool:
profiles/sanitizer_r02/README.md).  ``torch.empty`` is patched so that every CUDA buffer the host code allocates -
outputs, intermediates, workspaces - sits between two 4 KiB guard bands filled with a canary byte; after the launches
the bands must be untouched.  Shapes are chosen ragged on purpose (T not a multiple of 8 / 128, W not a multiple of 8,
channel counts that are not tile multiples, n not a multiple of 128) so every partial tile path runs."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from mqgan_b200 import ops, spec as S  # noqa: E402
from mqgan_b200.preencoder import PreEncoder  # noqa: E402
from mqgan_b200.synth import synth_state_dict, synth_mels  # noqa: E402

GUARD = 4096
CANARY = 0xA5


class GuardedAlloc:
    def __enter__(self):
        self.orig = torch.empty
        self.bufs = []
        torch.empty = self._empty
        return self

    def __exit__(self, *exc):
        torch.empty = self.orig
        return False

    def _empty(self, *shape, dtype=None, device=None, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        dev = torch.device(device) if device is not None else None
        if dev is None or dev.type != "cuda" or kw.get("pin_memory"):
            return self.orig(tuple(shape), dtype=dtype, device=device, **kw)
        dtype = dtype or torch.get_default_dtype()
        n = int(math.prod(shape)) * torch.zeros(1, dtype=dtype).element_size()
        pad = (-n) % 16
        raw = self.orig(n + pad + 2 * GUARD, dtype=torch.uint8, device=dev)
        raw.fill_(CANARY)
        self.bufs.append((raw, n))
        return raw[GUARD:GUARD + n].view(dtype).view(*shape)

    def check(self):
        torch.cuda.synchronize()
        bad = []
        for i, (raw, n) in enumerate(self.bufs):
            lo, hi = raw[:GUARD], raw[GUARD + n + ((-n) % 16):]
            if not bool((lo == CANARY).all()) or not bool((hi == CANARY).all()):
                bad.append((i, n, int((lo != CANARY).sum()), int((hi != CANARY).sum())))
        assert not bad, f"guard bands overwritten (buffer #, bytes, low, high): {bad[:8]} of {len(self.bufs)} buffers"
        return len(self.bufs)


def _rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("decoder_precision", ["bf16", "f16x2"])
@pytest.mark.parametrize("cfg_name,B,T,lengths", [
    ("TINY", 3, 277, [277, 130, 9]),        # T % 8 = 5, pair kernels on the 1-D (T >= 256) and 3x3 layers, ragged
    ("TINY", 2, 37, [37, 20]),              # small: tap-loop / halo kernels
    ("ODD", 2, 301, [301, 77]),             # 48-channel 1-D blocks, refiner widths 24 / 48 / 96 / 192, image width 72
])
def test_encode_decode_write_only_inside_their_buffers(cfg_name, B, T, lengths, decoder_precision):
    # "ODD": channel counts that are not multiples of 32 / 64 anywhere (the engine needs mel + mel/8 to be a multiple of 4)
    cfg = S.PreEncoderConfig(64, (48, 48, 64, 64), (3, 3, 5, 7), (8, 5, 5, 5), 24, 3, 8) if cfg_name == "ODD" else getattr(S, cfg_name)
    sd = synth_state_dict(cfg, 0)
    model = PreEncoder(cfg.mel_channels, list(cfg.channels), list(cfg.kernel_sizes), fsq_levels=list(cfg.fsq_levels),
                       dropout=0.0, refiner_base_channels=cfg.refiner_base_channels, refiner_depth=cfg.refiner_depth,
                       refiner_hidden_proj_divisor=cfg.refiner_hidden_proj_divisor, decoder_precision=decoder_precision)
    model.load_state_dict(sd)
    model = model.to("cuda").eval()
    mel = synth_mels(B, T, cfg.mel_channels, seed=3).cuda()
    pad = (torch.arange(T)[None, :] >= torch.tensor(lengths)[:, None]).unsqueeze(1).cuda()
    eng = model.engine()
    with GuardedAlloc() as g:
        idx = eng.encode(mel, pad)
        out = eng.decode(idx, pad)
        n = g.check()
    assert n > 40 and bool(torch.isfinite(out).all())
    with GuardedAlloc() as g:                 # no mask, and the exact-sum ConvBlock2D kernel
        eng.pre.table = eng.post.table = None
        out2 = eng.decode(eng.encode(mel, None), None)
        g.check()
    assert bool(torch.isfinite(out2).all())


def test_conv_vq_wgrad_mel_write_only_inside_their_buffers():
    dev = "cuda"
    with GuardedAlloc() as g:
        # every main loop x epilogue on shapes with partial tiles in H, W and Cout
        for kind, N, H, W, Cin, Cout, tail, kw in [
            ("linear", 2, 77, 1, 128, 144, (), {}),
            ("conv2d3", 1, 37, 20, 64, 96, (3, 3), {"pair": True}),
            ("conv2d3", 1, 37, 20, 64, 96, (3, 3), {"pair": False, "halo": True}),
            ("conv2d3", 1, 7, 36, 96, 72, (3, 3), {"pair": False, "halo": False}),
            ("causal1d", 1, 391, 1, 128, 160, (5,), {"pair": True}),
        ]:
            x = _rnd(N, H, W, Cin, seed=1).to(torch.bfloat16).to(dev)
            w = _rnd(Cout, Cin, *tail, seed=2) / (Cin * max(1, int(np.prod(tail)))) ** 0.5
            pc = ops.pack_conv(w, _rnd(Cout, seed=3), kind).to(dev)
            mask = (torch.arange(N * H) % 3 == 0).to(torch.uint8).to(dev)
            of = torch.empty(N, H, W, Cout, dtype=torch.float32, device=dev)
            ob = torch.empty(N, H, W, Cout, dtype=torch.bfloat16, device=dev)
            ops.conv_gemm(x, pc, N, H, W, out_f32=of, out_bf16=ob, act=True, row_mask=mask, mask_post=True, **kw)
            if Cout % 32 == 0:                                       # lean epilogue (bf16 only), staged stores
                ob2 = torch.empty(N, H, W, Cout, dtype=torch.bfloat16, device=dev)
                ops.conv_gemm(x, pc, N, H, W, out_bf16=ob2, act=True, row_mask=mask, mask_post=True, **kw)
        # fused pool epilogue on an image whose width is not a multiple of 8
        x = _rnd(1, 34, 20, 64, seed=4).to(torch.bfloat16).to(dev)
        pc = ops.pack_conv(_rnd(64, 64, 3, 3, seed=5) / 24.0, _rnd(64, seed=6), "conv2d3").to(dev)
        y = torch.empty(1, 34, 20, 64, dtype=torch.bfloat16, device=dev)
        yp = torch.empty(1, 17, 20, 64, dtype=torch.bfloat16, device=dev)
        m = torch.zeros(34, dtype=torch.uint8, device=dev)
        ops.conv_gemm(x, pc, 1, 34, 20, act=True, row_mask=m, mask_post=True, out_bf16=y, out_pool=yp, pair=True)
        # nearest-codeword lookup, n not a multiple of the 128-row tile
        for K, D, prec in ((1000, 4, "f16x2"), (300, 5, "bf16"), (1024, 64, "f16x2")):
            pcb = ops.pack_codebook(_rnd(K, D, seed=7), prec).to(dev)
            ops.vq_nearest(_rnd(1001, D, seed=8).to(dev), pcb, want_codes=True, want_dist=True)
        # weight gradient
        for kind, N, H, W, Cin, Cout, tail in (("conv2d3", 1, 13, 20, 64, 72, (3, 3)), ("causal1d", 1, 201, 1, 72, 64, (5,))):
            dh, dw = ops.conv_taps(kind, (Cout, Cin) + tail)
            ops.conv_wgrad(_rnd(N, H, W, Cout, seed=9).to(torch.bfloat16).to(dev), _rnd(N, H, W, Cin, seed=10).to(torch.bfloat16).to(dev),
                           N, H, W, Cout, Cin, dh, dw, split=3)
        from mqgan_b200.melspec import LogMelExtractor
        ext = LogMelExtractor({"sampling_rate": 44100, "filter_length": 2048, "hop_length": 512, "win_length": 2048,
                               "n_mel_channels": 128, "mel_fmin": 0.0, "mel_fmax": 22050.0}, dev)
        ext(_rnd(3, 9001, seed=11).to(dev), [9001, 5000, 1025])
        n = g.check()
    assert n > 20
