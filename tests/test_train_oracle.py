"""The training-step oracle (oracle/train_oracle.py) against outputs of the reference itself
(tests/golden/train_tiny.npz, written by oracle/make_golden_train.py from the reference's own PreEncoder,
discriminators, losses and Trainer step methods).  CPU only."""
import os

import numpy as np
import torch

from mqgan_b200 import spec as S
from mqgan_b200.synth import synth_disc_state_dict, synth_lengths, synth_mels, synth_state_dict
from oracle import train_oracle as TO

import pytest

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = ["train_tiny", "train_tiny_m"]


def tiny_train_state(name="train_tiny"):
    fx = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg, pdc, mbc = (getattr(S, str(n)) for n in fx["configs"])
    seed = int(fx["seed"])
    g_sd = synth_state_dict(cfg, seed=seed)
    g_sd["q_in_proj.weight"] = torch.from_numpy(fx["qin_w"]).clone()      # calibrated on the reference's latents
    g_sd["q_in_proj.bias"] = torch.from_numpy(fx["qin_b"]).clone()
    pd_sd = synth_disc_state_dict(S.patch_disc_param_spec(pdc), seed=seed)
    mb_sd = synth_disc_state_dict(S.multibin_param_spec(mbc), seed=seed + 1)
    return cfg, pdc, mbc, g_sd, pd_sd, mb_sd


def tiny_batch(step, B, T, n_mels):
    real = synth_mels(B, T, n_mels, seed=40 + step)
    lens = synth_lengths(B, T, seed=40 + step, ragged=True)
    return real.masked_fill((torch.arange(T)[None, :] >= lens[:, None]).unsqueeze(-1), 0.0), lens


@pytest.mark.parametrize("name", FIXTURES)
def test_train_oracle_matches_reference_two_iterations(name):
    fx = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg, pdc, mbc, g_sd, pd_sd, mb_sd = tiny_train_state(name)
    st = TO.TrainState(cfg, g_sd, pd_sd, mb_sd, TO.patch_cfg([k[0] for k in pdc.kernels], pdc.strides),
                       TO.multibin_cfg(mbc.kernel_sizes, mbc.n_bins, mbc.n_no_strides), dict(S.TINY_TRAIN))
    B, T = int(fx["B"]), int(fx["T"])
    g_keys = [str(k) for k in fx["g_keys"]]
    assert g_keys == list(g_sd)
    for step in (1, 2):
        real, lens = tiny_batch(step, B, T, cfg.mel_channels)
        o = TO.train_iteration(st, real, lens, gan=True, use_fm=step == 2)
        pre = f"s{step}_"
        got = np.array([o["loss_d"], o["loss_g_total"], o["loss_recon_pre"], o["loss_recon_post"], o["loss_gan"], o["loss_fm"]])
        # same library arithmetic in the same order: equal to float32 round-off
        np.testing.assert_allclose(got, fx[pre + "losses"], rtol=2e-6, atol=1e-7)
        assert float((o["recon_post"] - torch.from_numpy(fx[pre + "recon_post"])).abs().max()) < 1e-5
        assert float((o["recon_pre"] - torch.from_numpy(fx[pre + "recon_pre"])).abs().max()) < 1e-5
        gn = np.array([float(st.g[k].grad.norm()) if st.g[k].grad is not None else -1.0 for k in g_keys])
        ref = fx[pre + "g_grad_norms"]
        assert ((gn < 0) == (ref < 0)).all()            # hidden_proj never receives a gradient (detached refiner input)
        np.testing.assert_allclose(gn, ref, rtol=1e-4, atol=1e-9)
        for name in fx.files:
            if name.startswith(pre + "gg:"):
                k = name[len(pre) + 3:]
                g = st.g[k].grad
                r = torch.from_numpy(fx[name])
                assert float((g - r).abs().max()) <= 1e-5 * max(1e-6, float(r.abs().max())) + 1e-9, k
        ps = np.array([float(st.g[k].detach().double().sum()) for k in g_keys])
        np.testing.assert_allclose(ps, fx[pre + "g_param_sums"], rtol=1e-6, atol=1e-5)
        assert float((st.pd["convs.1.weight_u"] - torch.from_numpy(fx[pre + "d_u0"])).abs().max()) < 1e-6
        np.testing.assert_allclose([float(st.lecam.ema_real), float(st.lecam.ema_fake)], fx[pre + "lecam"], rtol=1e-5, atol=1e-8)
        d_now = {**{"pd:" + k: v for k, v in st.pd.items()}, **{"mb:" + k: v for k, v in st.mb.items()}}
        dps = np.array([float(d_now[str(k)].detach().double().sum()) for k in fx["d_keys"]])
        np.testing.assert_allclose(dps, fx[pre + "d_param_sums"], rtol=1e-6, atol=1e-5)


@pytest.mark.parametrize("name", FIXTURES)
def test_train_oracle_recon_only_iteration_after_two_gan_iterations(name):
    """The third pinned iteration runs before the GAN phase (epoch < discriminator_train_start_epoch): no discriminator
    step, no adversarial terms (train.py:526-527, 447-450)."""
    fx = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg, pdc, mbc, g_sd, pd_sd, mb_sd = tiny_train_state(name)
    st = TO.TrainState(cfg, g_sd, pd_sd, mb_sd, TO.patch_cfg([k[0] for k in pdc.kernels], pdc.strides),
                       TO.multibin_cfg(mbc.kernel_sizes, mbc.n_bins, mbc.n_no_strides), dict(S.TINY_TRAIN))
    B, T = int(fx["B"]), int(fx["T"])
    for step in (1, 2):
        real, lens = tiny_batch(step, B, T, cfg.mel_channels)
        TO.train_iteration(st, real, lens, gan=True, use_fm=step == 2)
    d_before = {k: v.detach().clone() for k, v in st.pd.items()}
    real, lens = tiny_batch(3, B, T, cfg.mel_channels)
    o = TO.train_iteration(st, real, lens, gan=False)
    got = np.array([o["loss_d"], o["loss_g_total"], o["loss_recon_pre"], o["loss_recon_post"], o["loss_gan"], o["loss_fm"]])
    np.testing.assert_allclose(got, fx["s3_losses"], rtol=2e-6, atol=1e-7)
    g_keys = [str(k) for k in fx["g_keys"]]
    gn = np.array([float(st.g[k].grad.norm()) if st.g[k].grad is not None else -1.0 for k in g_keys])
    np.testing.assert_allclose(gn, fx["s3_g_grad_norms"], rtol=1e-4, atol=1e-9)
    ps = np.array([float(st.g[k].detach().double().sum()) for k in g_keys])
    np.testing.assert_allclose(ps, fx["s3_g_param_sums"], rtol=1e-6, atol=1e-5)
    assert all(torch.equal(st.pd[k].detach(), d_before[k]) for k in d_before)          # the discriminators did not move


def test_hidden_proj_gets_no_gradient_and_refiner_input_is_detached():
    """preencoder.py:411-413: only the refiner sees x_post's gradient through the residual."""
    cfg, _, _, g_sd, _, _ = tiny_train_state()
    params = {k: v.clone().requires_grad_(True) for k, v in g_sd.items()}
    real, lens = tiny_batch(1, 2, 24, cfg.mel_channels)
    x_recon, x_post = TO.generator_forward(params, cfg, real, lens)
    (x_post - x_recon).sum().backward()                  # the residual alone
    assert params["hidden_proj.weight"].grad is None
    g = params["out_proj.weight"].grad                    # x_recon enters with +1 and -1: exactly zero
    assert g is None or float(g.abs().max()) == 0.0
    assert params["refiner.mid.conv1.parametrizations.weight.original1"].grad.abs().max() > 0
