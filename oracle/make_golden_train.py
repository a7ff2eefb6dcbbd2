"""Generate tests/golden/train_tiny.npz by running the UNMODIFIED reference training step here.

Run from the repo root, in the build container only (needs /root/reference):

    python oracle/make_golden_train.py

The reference's own modules are used as they are: ``preencoder.PreEncoder`` (train mode, dropout 0),
``discriminators.MelSpectrogramPatchDiscriminator2D`` / ``MultiBinDiscriminator``, ``losses.LSGANLoss`` /
``MaskedMelLoss`` and the two step functions ``train.Trainer._train_discriminator`` /
``_train_generator`` (train.py:380-501), which are called UNBOUND on a stand-in object that carries the
attributes ``Trainer.__init__`` would have set (train.py:202-227) - constructing a real ``Trainer`` needs a
dataset directory and a wandb session.  ``einx`` (FSQ training branch), ``matplotlib`` and ``wandb`` are
absent / unwanted here and are stubbed before import; none of them touches the arithmetic.

Weights and inputs are regenerated anywhere from (config, seed) by ``mqgan_b200.synth``; the fixture stores
only the reference's OUTPUTS for two consecutive iterations (the second with feature matching on):
losses, reconstructions, per-parameter gradient norms (as the step leaves them: clipped), a few complete
gradient tensors, per-parameter checksums of the updated weights and the spectral-norm ``u`` vectors.
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

FULL_GRADS_ALL = ["proj.weight", "pre.pw.parametrizations.weight.original0", "pre.dw.parametrizations.weight.original1",
              "encoder_blocks.0.conv1.parametrizations.weight.original1", "encoder_blocks.2.relu.beta",
              "encoder_blocks.1.cbam.spatial_attention.conv.weight", "encoder_blocks.1.residual.weight",
              "decoder_blocks.1.residual.weight", "q_in_proj.weight", "q_out_proj.weight",
              "decoder_blocks.2.conv2.weight_v", "decoder_blocks.0.conv1.weight_g", "out_proj.weight",
              "refiner.pre.conv2.parametrizations.weight.original1", "refiner.downs.0.conv.conv1.parametrizations.weight.original1",
              "refiner.ups.2.conv.conv1.parametrizations.weight.original0", "refiner.post.parametrizations.weight.original1",
              "refiner.reproj.weight"]
FULL_D_GRADS = ["pd:convs.0.weight_orig", "pd:convs.2.weight_orig", "pd:se_block.fc1.weight",
                "mb:discriminators.1.convs.1.weight_orig", "mb:discriminators.0.convs.3.bias"]


def import_reference():
    einx = types.ModuleType("einx")
    einx.where = lambda pattern, cond, a, b: torch.where(cond.view(-1, *([1] * (a.dim() - 1))), a, b)
    sys.modules.setdefault("einx", einx)
    for name in ("wandb", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import train as ref_train  # noqa
    return ref_train


CASES = {
    # fixture name: (generator config, patch D, multi-bin D, B, T, seed)
    "train_tiny": ("TINY", "TINY_PATCH_D", "TINY_MULTIBIN_D", 4, 48, 3),
    "train_tiny_m": ("TINY_M", "TINY_M_PATCH_D", "TINY_M_MULTIBIN_D", 3, 43, 5),     # T not a multiple of 8
}


def main():
    T_ = import_reference()
    for name, case in CASES.items():
        run_case(T_, name, *case)


def run_case(T_, name, cfg_name, pd_name, mb_name, B, T, seed):
    from mqgan_b200 import spec as S
    from mqgan_b200.synth import synth_state_dict, synth_mels, synth_lengths, synth_disc_state_dict
    from oracle import train_oracle as TO

    torch.set_num_threads(os.cpu_count() or 1)
    cfg, pdc, mbc, tcfg = getattr(S, cfg_name), getattr(S, pd_name), getattr(S, mb_name), S.TINY_TRAIN
    g_sd = synth_state_dict(cfg, seed=seed)
    pd_sd = synth_disc_state_dict(S.patch_disc_param_spec(pdc), seed=seed)
    mb_sd = synth_disc_state_dict(S.multibin_param_spec(mbc), seed=seed + 1)

    # with default init every frame quantises to one code (all codes 0 -> q_out_proj sees no gradient);
    # recalibrate q_in_proj on the reference's own latents so the codes vary.  Stored in the fixture.
    from oracle.make_golden import build_reference_model, reference_latents
    from mqgan_b200.synth import recalibrate_q_in_proj
    import preencoder as ref_pre
    z0, _ = reference_latents(build_reference_model(ref_pre, cfg, g_sd), synth_mels(4, 96, cfg.mel_channels, seed=100), None)
    recalibrate_q_in_proj(g_sd, z0)

    gen = T_.MVQGenerator(cfg.mel_channels, list(cfg.channels), list(cfg.kernel_sizes), fsq_levels=list(cfg.fsq_levels),
                          dropout=0.0, refiner_base_channels=cfg.refiner_base_channels, refiner_depth=cfg.refiner_depth,
                          refiner_hidden_proj_divisor=cfg.refiner_hidden_proj_divisor)
    gen.load_state_dict(g_sd, strict=True)
    # dropout = 0.0 does not reach every layer: DownBlock / UpBlock build their ConvBlock with the default 0.1
    # (preencoder.py:109, 121) and so do `pre` / `post` (ConvBlock2D default, :233, :322, :351).  Dropout draws
    # from torch's global RNG stream, which no other implementation can reproduce, so the golden run sets the
    # PROBABILITY of every nn.Dropout to 0 (an attribute, not code).
    n_do = 0
    for m in gen.modules():
        if isinstance(m, torch.nn.Dropout) and m.p != 0.0:
            m.p = 0.0
            n_do += 1
    print(f"set p = 0 on {n_do} hard-wired Dropout layers")
    pd = T_.MelSpectrogramPatchDiscriminator2D(pdc.mel_channels, hidden_channels=list(pdc.hidden_channels),
                                               kernel_sizes=[k[0] for k in pdc.kernels], stride=[list(s) for s in pdc.strides])
    pd.load_state_dict(pd_sd, strict=True)
    mb = T_.MultiBinDiscriminator(mbc.mel_channels, hidden_channels=list(mbc.hidden_channels),
                                  kernel_sizes=list(mbc.kernel_sizes), n_bins=mbc.n_bins, n_no_strides=mbc.n_no_strides)
    mb.load_state_dict(mb_sd, strict=True)

    me = types.SimpleNamespace()
    me.config = {"training": dict(tcfg)}
    me.device = torch.device("cpu")
    me.generator, me.patch_discriminator, me.multibin_discriminator = gen, pd, mb
    me.optimizer_g, me.optimizer_d = T_.Trainer._init_optimizers(me)                 # train.py:312-324
    me.scheduler_g = T_.Trainer._init_schedulers(me)                                  # :326-329
    me.gan_loss = T_.LSGANLoss()
    me.recon_loss_all = T_.MaskedMelLoss("mse")
    me.recon_loss_group = T_.MaskedMelLoss("mse", group_size=16)
    me.scaler_g = T_.GradScaler(enabled=False)
    me.scaler_d = T_.GradScaler(enabled=False)
    me.epoch = tcfg["discriminator_train_start_epoch"]                                # GAN terms active

    # the oracle runs beside it from the same state
    st = TO.TrainState(cfg, g_sd, pd_sd, mb_sd, TO.patch_cfg([k[0] for k in pdc.kernels], pdc.strides),
                       TO.multibin_cfg(mbc.kernel_sizes, mbc.n_bins, mbc.n_no_strides), dict(tcfg))

    gen.train(); pd.train(); mb.train()                                               # train.py:504-506
    out = {}
    for step in (1, 2):
        me.config["training"]["use_fm_loss"] = step == 2
        real = synth_mels(B, T, cfg.mel_channels, seed=40 + step)
        lens = synth_lengths(B, T, seed=40 + step, ragged=True)
        real = real.masked_fill((torch.arange(T)[None, :] >= lens[:, None]).unsqueeze(-1), 0.0)
        recon_pre, recon_post = gen(real, lens)                                       # :524
        loss_d = T_.Trainer._train_discriminator(me, real, recon_post, lens)          # :528
        d_named = [("pd:" + k, p) for k, p in pd.named_parameters()] + [("mb:" + k, p) for k, p in mb.named_parameters()]
        d_gn = {k: float(p.grad.norm()) for k, p in d_named}
        d_full = {k: p.grad.detach().clone() for k, p in d_named if k in FULL_D_GRADS}
        g_losses = T_.Trainer._train_generator(me, real, recon_pre, recon_post, lens)  # :530
        g_named = dict(gen.named_parameters())
        # legacy weight-norm exposes weight_g / weight_v as the parameters; names match the state-dict
        g_gn = {k: (float(g_named[k].grad.norm()) if g_named[k].grad is not None else -1.0) for k in g_sd}   # -1: no gradient
        pre = f"s{step}_"
        out[pre + "losses"] = np.array([loss_d, g_losses["loss_g_total"], g_losses["loss_recon_pre"],
                                        g_losses["loss_recon_post"], g_losses["loss_gan"], g_losses["loss_fm"]], np.float64)
        out[pre + "recon_pre"] = recon_pre.detach().numpy()
        out[pre + "recon_post"] = recon_post.detach().numpy()
        out[pre + "g_grad_norms"] = np.array([g_gn[k] for k in g_sd], np.float64)
        out[pre + "d_grad_keys"] = np.array([k for k, _ in d_named])
        out[pre + "d_grad_norms"] = np.array([d_gn[k] for k, _ in d_named], np.float64)
        for k in FULL_GRADS_ALL:
            if k in g_named and g_named[k].grad is not None:
                out[pre + "gg:" + k] = g_named[k].grad.detach().numpy().copy()
        for k, v in d_full.items():
            out[pre + "dg:" + k] = v.numpy()
        gsd_now = gen.state_dict()
        out[pre + "g_param_sums"] = np.array([float(gsd_now[k].double().sum()) for k in g_sd], np.float64)
        out[pre + "g_param_delta"] = np.array([float((gsd_now[k] - g_sd[k]).double().abs().sum()) for k in g_sd], np.float64)
        dsd_now = {**{"pd:" + k: v for k, v in pd.state_dict().items()}, **{"mb:" + k: v for k, v in mb.state_dict().items()}}
        d_all_keys = ["pd:" + k for k in pd_sd] + ["mb:" + k for k in mb_sd]
        out[pre + "d_param_sums"] = np.array([float(dsd_now[k].double().sum()) for k in d_all_keys], np.float64)
        out[pre + "d_u0"] = dsd_now["pd:convs.1.weight_u"].numpy().copy()
        out[pre + "lecam"] = np.array([float(me.gan_loss.ema_real), float(me.gan_loss.ema_fake)], np.float64)
        out[pre + "lr_g"] = np.array(me.scheduler_g.get_last_lr()[0])

        # ---- the restatement, same batch ----
        o = TO.train_iteration(st, real, lens, gan=True, use_fm=step == 2)
        ol = np.array([o["loss_d"], o["loss_g_total"], o["loss_recon_pre"], o["loss_recon_post"], o["loss_gan"], o["loss_fm"]])
        gn_o = np.array([float(st.g[k].grad.norm()) if st.g[k].grad is not None else -1.0 for k in g_sd])
        rel = np.abs(gn_o - out[pre + "g_grad_norms"]) / (np.abs(out[pre + "g_grad_norms"]) + 1e-12)
        print(f"[step {step}] reference losses {out[pre + 'losses']}")
        print(f"[step {step}] oracle    losses {ol}")
        print(f"[step {step}] |recon_post diff|max {float((o['recon_post'] - recon_post.detach()).abs().max()):.3e}  "
              f"G grad-norm rel diff max {rel.max():.3e}  lecam ref {out[pre + 'lecam']} oracle "
              f"{[float(st.lecam.ema_real), float(st.lecam.ema_fake)]}")
        ps_o = np.array([float(st.g[k].detach().double().sum()) for k in g_sd])
        print(f"[step {step}] G param-sum diff max {np.abs(ps_o - out[pre + 'g_param_sums']).max():.3e}; "
              f"u diff {float((st.pd['convs.1.weight_u'] - dsd_now['pd:convs.1.weight_u']).abs().max()):.3e}")

    # ---- a third iteration before the GAN phase (epoch < discriminator_train_start_epoch): no discriminator step, no
    # adversarial / feature-matching terms (train.py:526-527, 447-450) ----
    me.epoch = tcfg["discriminator_train_start_epoch"] - 1
    me.config["training"]["use_fm_loss"] = False
    real = synth_mels(B, T, cfg.mel_channels, seed=43)
    lens = synth_lengths(B, T, seed=43, ragged=True)
    real = real.masked_fill((torch.arange(T)[None, :] >= lens[:, None]).unsqueeze(-1), 0.0)
    recon_pre, recon_post = gen(real, lens)
    g_losses = T_.Trainer._train_generator(me, real, recon_pre, recon_post, lens)
    g_named = dict(gen.named_parameters())
    out["s3_losses"] = np.array([0.0, g_losses["loss_g_total"], g_losses["loss_recon_pre"], g_losses["loss_recon_post"],
                                 g_losses["loss_gan"], g_losses["loss_fm"]], np.float64)
    out["s3_g_grad_norms"] = np.array([float(g_named[k].grad.norm()) if g_named[k].grad is not None else -1.0 for k in g_sd],
                                      np.float64)
    gsd_now = gen.state_dict()
    out["s3_g_param_sums"] = np.array([float(gsd_now[k].double().sum()) for k in g_sd], np.float64)
    o = TO.train_iteration(st, real, lens, gan=False)
    print(f"[step 3, recon only] reference {out['s3_losses']}  oracle "
          f"{[o['loss_d'], o['loss_g_total'], o['loss_recon_pre'], o['loss_recon_post'], o['loss_gan'], o['loss_fm']]}")

    out["B"], out["T"], out["seed"] = np.array(B), np.array(T), np.array(seed)
    out["configs"] = np.array([cfg_name, pd_name, mb_name])
    out["qin_w"], out["qin_b"] = g_sd["q_in_proj.weight"].numpy(), g_sd["q_in_proj.bias"].numpy()
    out["g_keys"] = np.array(list(g_sd))
    out["d_keys"] = np.array(["pd:" + k for k in pd_sd] + ["mb:" + k for k in mb_sd])
    path = os.path.join(REPO, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
