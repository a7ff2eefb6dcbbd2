"""BASELINE-size checks (configs[1]: hifispeech, 256 x 1024 frames) through size-independent
properties, plus an oracle comparison on a slice that the CPU finishes in seconds."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from mqgan_b200 import spec as S  # noqa: E402
from mqgan_b200.synth import synth_mels  # noqa: E402
from oracle import preencoder_oracle as O  # noqa: E402
from tests.helpers import index_report  # noqa: E402


@pytest.fixture(scope="module")
def full():
    import bench
    cfg = S.HIFISPEECH
    model, sd = bench.build_model(cfg, "f16x2", torch.device("cuda"))
    B, T = 256, 1024
    mel = synth_mels(B, T, cfg.mel_channels, seed=1)
    lengths = torch.randint(256, T + 1, (B,), generator=torch.Generator().manual_seed(3))
    lengths[0] = T
    pad = torch.arange(T)[None, :] >= lengths[:, None]
    mel = mel.masked_fill(pad.unsqueeze(-1), 0.0)
    return cfg, sd, model, mel.cuda(), pad.unsqueeze(1).cuda(), lengths


def test_full_batch_encode_decode_properties(full):
    cfg, sd, model, mel, mask, lengths = full
    idx = model.encode(mel, mask)
    assert idx.shape == (256, 1024) and idx.dtype == torch.int64
    assert int(idx.min()) >= 0 and int(idx.max()) < cfg.codebook_size
    assert len(torch.unique(idx)) > 500                      # the calibrated model uses most codes
    # determinism (no float atomics anywhere on the path)
    assert torch.equal(idx, model.encode(mel, mask))
    # utterances only interact through the padded length: any sub-batch alone gives identical indices
    sel = torch.tensor([0, 17, 100, 255])
    assert torch.equal(idx[sel], model.encode(mel[sel], mask[sel]))
    out = model.decode(idx, mask)
    assert out.shape == mel.shape and bool(torch.isfinite(out).all())
    assert torch.equal(out[sel], model.decode(idx[sel], mask[sel]))      # decoder + refiner likewise
    # decoder + refiner are padding-invariant: trimming the padding of a short
    # utterance does not change its valid frames (up to the 8-frame refiner alignment)
    b = int(torch.argmin(lengths))
    L = int(lengths[b]) // 8 * 8
    alone = model.decode(idx[b:b + 1, :L], mask[b:b + 1, :, :L])
    assert float((alone[0, :L - 64] - out[b, :L - 64]).abs().max()) < 1e-5
    # FSQ is idempotent on its own codes: decode's gather input re-quantises to the same index
    eng = model.engine()
    codes = O.fsq_indices_to_codes(torch.arange(cfg.codebook_size), cfg.fsq_levels)
    from mqgan_b200 import ops
    # codes are already in bounded space; atanh maps them back to pre-bound latents
    lv, basis, half_l, offset, shift, half_w = O.fsq_constants(cfg.fsq_levels)
    z = torch.atanh(((codes * half_w) + offset) / half_l) - shift
    re_idx = ops.fsq_quantize(z.float().cuda().contiguous(), eng.fsq)
    assert torch.equal(re_idx.cpu(), torch.arange(cfg.codebook_size))


def test_full_batch_slice_vs_oracle(full):
    cfg, sd, model, mel, mask, lengths = full
    sel = [0, 5]                                             # one full-length, one ragged utterance
    idx = model.encode(mel, mask)[sel].cpu()
    w = O.effective_weights(sd)
    m = mask[sel].cpu()
    x = mel[sel].cpu()
    z32 = O.encode_latents(w, cfg, x, m, folded=True)
    ref_idx = O.fsq_quantize(z32, cfg.fsq_levels)[1]
    rep = index_report(idx, ref_idx)
    # margin-aware gate without the (slow) float64 pass: disagreeing frames must sit within 2e-4 of a
    # rounding boundary in the oracle's own fp32 latents
    margin = O.fsq_round_margin(z32, cfg.fsq_levels)
    bad = (idx != ref_idx) & (margin > 2e-4)
    print(rep)
    assert int(bad.sum()) == 0 and rep["agree"] >= 0.995, rep
    ref = O.decode(w, cfg, ref_idx, m, folded=True)
    out = model.decode(ref_idx.cuda(), m.cuda()).cpu()
    assert float((out - ref).abs().max()) <= 2e-2 + 2e-2 * float(ref.abs().max())
