"""One small launch of every kernel family, to be run under compute-sanitizer (tools/sanitize.sh):

    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitizer_cases.py

Covers conv_gemm_kernel (tap loop), conv_halo_kernel, conv_pair_kernel (bf16, fused upsample-concat, three-segment
f16x2), conv_pair1d_kernel, the fused pool epilogue, vq_nearest_kernel (bf16 + f16x2, narrow and wide codebooks),
conv_wgrad, log_mel, and through a tiny end-to-end encode + decode in both decoder precisions every pointwise kernel
of the inference path (convblock2d table + exact, CBAM, q_in + FSQ, gather, refiner masks / stem / pools / upcat /
tail).  Prints a checksum per case so a silent no-op would be visible; the numeric checks live in tests/.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mqgan_b200 import ops, spec as S
from mqgan_b200.preencoder import PreEncoder
from mqgan_b200.synth import synth_state_dict, synth_mels

DEV = "cuda"


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def conv_case(tag, kind, N, H, W, Cin, Cout, tail, split=False, **kw):
    x = rnd(N, H, W, Cin, seed=1)
    w = rnd(Cout, Cin, *tail, seed=2) / (Cin * max(1, int(np.prod(tail)))) ** 0.5
    pc = ops.pack_conv(w, rnd(Cout, seed=3), kind, split=split).to(DEV)
    xin = ops.split_bf16(x.reshape(-1, Cin).to(DEV), 2) if split == "f16x2" else x.to(torch.bfloat16).to(DEV)
    out = torch.empty(N, H, W, Cout, dtype=torch.float32, device=DEV)
    ops.conv_gemm(xin, pc, N, H, W, out_f32=out, **kw)
    torch.cuda.synchronize()
    print(tag, float(out.double().sum()))


def main():
    torch.cuda.set_device(0)
    conv_case("conv_gemm_kernel linear", "linear", 2, 77, 1, 128, 144, ())
    conv_case("conv_gemm_kernel taps 3x3", "conv2d3", 1, 7, 36, 96, 64, (3, 3), pair=False, halo=False)
    conv_case("conv_halo_kernel", "conv2d3", 1, 48, 24, 64, 64, (3, 3), pair=False, halo=True)
    conv_case("conv_pair_kernel bf16", "conv2d3", 1, 64, 24, 128, 128, (3, 3), pair=True)
    conv_case("conv_pair_kernel f16x2 3 segments", "conv2d3", 1, 64, 24, 64, 64, (3, 3), split="f16x2", pair=True)
    conv_case("conv_pair1d_kernel", "causal1d", 1, 512, 1, 128, 256, (5,), pair=True)
    conv_case("conv_pair1d_kernel f16x2", "same1d", 1, 300, 1, 64, 128, (3,), split="f16x2", pair=True)
    # fused upsample + concat, fused pool epilogue
    N, Hl, W, Cx, Cs, Cout = 1, 32, 16, 128, 64, 64
    wu = rnd(Cout, Cx + Cs, 3, 3, seed=5) / (9 * (Cx + Cs)) ** 0.5
    pcu = ops.pack_upconv(wu, rnd(Cout, seed=6), Cx, Cs).to(DEV)
    xl = rnd(N, Hl, W, Cx, seed=7).to(torch.bfloat16).to(DEV)
    sk = rnd(N, 2 * Hl, W, Cs, seed=8).to(torch.bfloat16).to(DEV)
    t = torch.empty(N, 2 * Hl, W, Cout, dtype=torch.bfloat16, device=DEV)
    ops.conv_gemm(xl, pcu, N, Hl, W, x2=sk, act=True, out_bf16=t, pair=True)
    yp = torch.empty(N, Hl, W, Cout, dtype=torch.bfloat16, device=DEV)
    pc2 = ops.pack_conv(rnd(Cout, Cout, 3, 3, seed=9) / 24.0, rnd(Cout, seed=10), "conv2d3").to(DEV)
    y = torch.empty_like(t)
    mask = (torch.arange(N * 2 * Hl) % 5 == 1).to(torch.uint8).to(DEV)
    ops.conv_gemm(t, pc2, N, 2 * Hl, W, act=True, row_mask=mask, mask_post=True, out_bf16=y, out_pool=yp, pair=True)
    torch.cuda.synchronize()
    print("conv_pair_kernel up-concat + pool epilogue", float(y.double().sum()), float(yp.double().sum()))
    # nearest-codeword lookup
    for K, D, prec in ((1024, 4, "f16x2"), (512, 5, "bf16"), (1024, 64, "f16x2"), (256, 24, "bf16")):
        pc = ops.pack_codebook(rnd(K, D, seed=11), prec).to(DEV)
        idx, codes = ops.vq_nearest(rnd(3000, D, seed=12).to(DEV), pc)
        torch.cuda.synchronize()
        print("vq_nearest_kernel", K, D, prec, int(idx.sum()))
    sink = torch.zeros(1, device=DEV)
    from mqgan_b200 import _lib
    _lib.call("mq_tmem_read_probe", 4, sink.data_ptr(), torch.cuda.current_stream().cuda_stream)
    # weight gradient
    for kind, N, H, W, Cin, Cout, tail in (("conv2d3", 1, 16, 24, 64, 64, (3, 3)), ("causal1d", 1, 200, 1, 128, 64, (5,))):
        dh, dw = ops.conv_taps(kind, (Cout, Cin) + tail)
        g = ops.conv_wgrad(rnd(N, H, W, Cout, seed=13).to(torch.bfloat16).to(DEV), rnd(N, H, W, Cin, seed=14).to(torch.bfloat16).to(DEV),
                           N, H, W, Cout, Cin, dh, dw, split=2)
        torch.cuda.synchronize()
        print("conv_wgrad", kind, float(g.double().sum()))
    # log-mel front-end
    from mqgan_b200.melspec import LogMelExtractor
    ext = LogMelExtractor({"sampling_rate": 44100, "filter_length": 2048, "hop_length": 512, "win_length": 2048,
                           "n_mel_channels": 128, "mel_fmin": 0.0, "mel_fmax": 22050.0}, DEV)
    m, fr = ext(rnd(2, 9000, seed=15).to(DEV), [9000, 5000])
    torch.cuda.synchronize()
    print("log_mel", float(m.double().sum()), fr)
    # end to end, both decoder precisions (T >= 256 so the 1-D layers take the pair kernel, T8 >= 32 the 3x3 pair kernel)
    cfg = S.TINY
    sd = synth_state_dict(cfg, 0)
    mel = synth_mels(2, 264, cfg.mel_channels, seed=1)
    lengths = torch.tensor([264, 150])
    pad = (torch.arange(264)[None, :] >= lengths[:, None])
    for dprec in ("bf16", "f16x2"):
        for table in (True, False):
            m_ = PreEncoder(cfg.mel_channels, list(cfg.channels), list(cfg.kernel_sizes), fsq_levels=list(cfg.fsq_levels),
                            dropout=0.0, refiner_base_channels=cfg.refiner_base_channels, refiner_depth=cfg.refiner_depth,
                            decoder_precision=dprec)
            m_.load_state_dict(sd)
            m_ = m_.to(DEV).eval()
            if not table:
                eng = m_.engine()
                eng.pre.table = eng.post.table = None          # exact-sum ConvBlock2D kernel
            idx = m_.encode(mel.to(DEV), pad.unsqueeze(1).to(DEV))
            out = m_.decode(idx, pad.unsqueeze(1).to(DEV))
            torch.cuda.synchronize()
            print("end to end", dprec, "table" if table else "exact", int(idx.sum()), float(out.double().sum()))
    print("sanitizer cases done")


if __name__ == "__main__":
    main()
