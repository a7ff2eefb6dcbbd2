"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

Run from the repo root, in the build container only (needs /root/reference):

    python oracle/make_golden.py

The reference (a Python package of bare top-level modules) is put on
sys.path read-only; the one missing dependency ``einx`` (used only in FSQ's
training branch, quantizer.py:151,160) is stubbed before import.  Nothing from
the reference is copied into the repo: only its numeric OUTPUTS on synthetic
weights/inputs, which ``mqgan_b200.synth`` regenerates anywhere from (config,
seed).  The calibrated ``q_in_proj`` tensors are stored in the
fixture because they depend on running the reference encoder.

Each fixture holds, for one config:
  qin_w, qin_b          calibrated q_in_proj (float32)
  lengths               (B,) int64 ragged lengths; T = mels.shape[1]
  z                     reference pre-quantiser latents (B,T,D) float32
  indices               reference encode() output (B,T) int16
  recon                 reference decode(indices) output (B,T,n_mels) float32
  rand_indices, rand_recon   decode() of random indices (decoder pinned alone)
  nomask_indices, nomask_recon  encode/decode with x_mask=None on utterance 0
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

REF = os.environ.get("MQGAN_REFERENCE", "/root/reference")


def import_reference():
    einx = types.ModuleType("einx")
    einx.where = lambda pattern, cond, a, b: torch.where(cond.view(-1, *([1] * (a.dim() - 1))), a, b)
    sys.modules.setdefault("einx", einx)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import preencoder as ref_pre  # noqa
    return ref_pre


def build_reference_model(ref_pre, cfg, sd):
    m = ref_pre.PreEncoder(cfg.mel_channels, list(cfg.channels), list(cfg.kernel_sizes),
                           fsq_levels=list(cfg.fsq_levels), dropout=0.0,
                           refiner_base_channels=cfg.refiner_base_channels,
                           refiner_depth=cfg.refiner_depth,
                           refiner_hidden_proj_divisor=cfg.refiner_hidden_proj_divisor)
    m.load_state_dict(sd, strict=True)
    return m.eval()


def reference_latents(model, mel, mask):
    grabbed = {}
    h = model.q_in_proj.register_forward_hook(lambda mod, inp, out: grabbed.__setitem__("z", out.detach()))
    with torch.no_grad():
        idx = model.encode(mel, mask)
    h.remove()
    return grabbed["z"], idx


CASES = {
    # name: (config attr, B, T, calibration (B, T))
    "tiny": ("TINY", 4, 53, (4, 96)),
    "hifispeech": ("HIFISPEECH", 3, 50, (4, 96)),
    "hifimusic": ("HIFIMUSIC", 2, 43, (4, 96)),
    # amplified weights (mqgan_b200.synth.amplify_state_dict): APTx saturated, ConvBlock2D's stencil value outside its
    # +-64 table, re-encoded mels in the log-mel range [-11.5, 5] - the magnitudes random-init never reaches
    "tiny_amp": ("TINY", 4, 53, (4, 96)),
    "hifispeech_amp": ("HIFISPEECH", 3, 50, (4, 96)),
}


def fsq_adversarial(levels):
    """Latents whose bounded image sits ON a rounding boundary k + 0.5 (ties: half to even, quantizer.py:137 via
    torch.round), their +-1..3 ulp neighbours, saturating |z| and zeros - the cases a random fixture never hits."""
    lv = torch.tensor(levels, dtype=torch.float64)
    half_l = (lv - 1) * (1 + 1e-3) / 2
    offset = torch.where(lv % 2 == 0, 0.5, 0.0).double()
    shift = torch.atanh(offset / half_l)
    rows = []
    D = len(levels)
    for d in range(D):
        L = levels[d]
        lo = -(L // 2)
        for k in range(lo, lo + L - 1):                      # boundaries between adjacent levels
            b = k + 0.5
            arg = (b + float(offset[d])) / float(half_l[d])
            if abs(arg) >= 1:
                continue
            z0 = torch.tensor(float(torch.atanh(torch.tensor(arg, dtype=torch.float64)) - shift[d]), dtype=torch.float32)
            cands = [z0]
            up, dn = z0.clone(), z0.clone()
            for _ in range(3):
                up = torch.nextafter(up, torch.tensor(float("inf")))
                dn = torch.nextafter(dn, torch.tensor(float("-inf")))
                cands += [up.clone(), dn.clone()]
            for c in cands:
                for base in (0.0, 0.3, -0.7):                # the other dims sit safely inside a level
                    row = torch.full((D,), base, dtype=torch.float32)
                    row[d] = c
                    rows.append(row)
    for big in (8.0, 20.0, 88.0, 1e4, 3e38):
        rows.append(torch.full((D,), big))
        rows.append(torch.full((D,), -big))
    rows.append(torch.zeros(D))
    rows.append(torch.full((D,), -0.0))
    return torch.stack(rows)


def main():
    from mqgan_b200 import spec as S
    from mqgan_b200.synth import synth_state_dict, synth_mels, synth_lengths, recalibrate_q_in_proj, amplify_state_dict
    from oracle import preencoder_oracle as O

    ref_pre = import_reference()
    torch.set_num_threads(os.cpu_count() or 1)
    out_dir = os.path.join(REPO, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

    for name, (cfg_name, B, T, (cb, ct)) in CASES.items():
        cfg = getattr(S, cfg_name)
        sd = synth_state_dict(cfg, seed=0)
        amplified = name.endswith("_amp")
        if amplified:
            sd = amplify_state_dict(sd)
        model = build_reference_model(ref_pre, cfg, sd)
        # --- SURVEY D4 calibration, with the reference's own encoder ---------
        cal = synth_mels(cb, ct, cfg.mel_channels, seed=100)
        z0, idx0 = reference_latents(model, cal, None)
        print(f"[{name}] uncalibrated: unique indices {idx0.unique().numel()}, z std {z0.std(dim=(0,1)).tolist()}")
        recalibrate_q_in_proj(sd, z0)
        model = build_reference_model(ref_pre, cfg, sd)
        # --- the pinned batch ----------------------------------------------
        mel = synth_mels(B, T, cfg.mel_channels, seed=1)
        lengths = synth_lengths(B, T, seed=1, ragged=True)
        mask = ref_pre.sequence_mask(T, lengths).unsqueeze(1)
        mel = mel.masked_fill(mask.squeeze(1).unsqueeze(-1), 0.0)   # CLI zero-pads (reencode_spectrograms.py:54-59)
        z, idx = reference_latents(model, mel, mask)
        with torch.no_grad():
            recon = model.decode(idx, mask)
            g = torch.Generator().manual_seed(5)
            rand_idx = torch.randint(0, cfg.codebook_size, (B, T), generator=g)
            rand_recon = model.decode(rand_idx, mask)
            nm_idx = model.encode(mel[:1], None)
            nm_recon = model.decode(nm_idx, None)
        print(f"[{name}] magnitudes: |recon| max {float(recon.abs().max()):.2f} range [{float(recon.min()):.2f}, {float(recon.max()):.2f}]")
        print(f"[{name}] calibrated: unique indices {idx.unique().numel()} of {idx.numel()} frames")
        # --- cross-check the restatement before committing ---------------------
        w = O.effective_weights(sd)
        z_o = O.encode_latents(w, cfg, mel, mask, folded=True)
        idx_o = O.fsq_quantize(z_o, cfg.fsq_levels)[1]
        recon_o = O.decode(w, cfg, idx, mask, folded=True)
        print(f"[{name}] oracle vs reference: |dz|max {float((z_o - z).abs().max()):.3e}  "
              f"idx mismatches {int((idx_o != idx).sum())}  |drecon|max {float((recon_o - recon).abs().max()):.3e}")
        np.savez_compressed(
            os.path.join(out_dir, f"{name}.npz"),
            config=np.array(cfg_name), seed=np.array(0), mel_seed=np.array(1), amplified=np.array(int(amplified)),
            qin_w=sd["q_in_proj.weight"].numpy(), qin_b=sd["q_in_proj.bias"].numpy(),
            lengths=lengths.numpy(), T=np.array(T),
            z=z.numpy(), indices=idx.numpy().astype(np.int16), recon=recon.numpy(),
            rand_indices=rand_idx.numpy().astype(np.int16), rand_recon=rand_recon.numpy(),
            nomask_indices=nm_idx.numpy().astype(np.int16), nomask_recon=nm_recon.numpy(),
        )
        print(f"[{name}] wrote {os.path.join(out_dir, name + '.npz')}")

    # FSQ known-answer table straight from the reference's quantizer (integer pin)
    import quantizer as ref_q
    for levels in ([8, 5, 5, 5], [8, 8, 5, 5, 5]):
        fsq = ref_q.FSQ(levels=levels).eval()
        g = torch.Generator().manual_seed(11)
        zz = torch.randn(1, 4096, len(levels), generator=g) * 1.5
        # adversarial: values whose bounded image sits on / next to k + 0.5, saturating magnitudes, zeros
        z_adv = fsq_adversarial(levels).unsqueeze(0)
        # 2^20 random latents, regenerated from the seed by the tests (only the reference's answers are stored)
        g2 = torch.Generator().manual_seed(12)
        z_big = torch.randn(1, 1 << 20, len(levels), generator=g2) * 1.5
        with torch.no_grad():
            codes, idx = fsq(zz)
            codes_adv, idx_adv = fsq(z_adv)
            _, idx_big = fsq(z_big)
            all_idx = torch.arange(int(np.prod(levels)))
            all_codes = fsq.indices_to_codes(all_idx)
        np.savez_compressed(os.path.join(out_dir, "fsq_" + "_".join(map(str, levels)) + ".npz"),
                            levels=np.array(levels), z=zz.numpy(), codes=codes.numpy(),
                            indices=idx.numpy().astype(np.int32), all_codes=all_codes.numpy(),
                            z_adv=z_adv[0].numpy(), codes_adv=codes_adv[0].numpy(), indices_adv=idx_adv[0].numpy().astype(np.int32),
                            big_seed=np.array(12), big_n=np.array(1 << 20), big_scale=np.array(1.5),
                            indices_big=idx_big[0].numpy().astype(np.uint16))
        print(f"[fsq {levels}] adversarial rows {z_adv.shape[1]}, big rows {z_big.shape[1]}")
        print(f"[fsq {levels}] wrote fixture")


if __name__ == "__main__":
    main()
