"""Training-step parity report on the GPU: mqgan_b200.training.TrainStep (TINY config, two iterations) against
tests/golden/train_tiny.npz (outputs of the reference's own Trainer step).  Prints one JSON object; the
thresholds in tests/test_gpu_training.py come from this report.  Usage: python tools/train_parity.py [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mqgan_b200 import spec as S
from mqgan_b200 import training as TR
from mqgan_b200.synth import synth_disc_state_dict, synth_lengths, synth_mels, synth_state_dict


def main():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    rep = {}
    for fixture in ("train_tiny", "train_tiny_m"):
        rep.update(run_fixture(fixture))
    print(json.dumps(rep, indent=1))
    if len(sys.argv) > 1:
        json.dump(rep, open(sys.argv[1], "w"), indent=1)


def run_fixture(fixture):
    fx = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", fixture + ".npz"))
    cfg, pdc, mbc = (getattr(S, str(n)) for n in fx["configs"])
    seed = int(fx["seed"])
    g_sd = synth_state_dict(cfg, seed=seed)
    g_sd["q_in_proj.weight"] = torch.from_numpy(fx["qin_w"]).clone()
    g_sd["q_in_proj.bias"] = torch.from_numpy(fx["qin_b"]).clone()
    pd_sd = synth_disc_state_dict(S.patch_disc_param_spec(pdc), seed=seed)
    mb_sd = synth_disc_state_dict(S.multibin_param_spec(mbc), seed=seed + 1)
    rep = {}
    for native in (True, False):
        ts = TR.TrainStep(cfg, pdc, mbc, g_sd, pd_sd, mb_sd, dict(S.TINY_TRAIN), "cuda", native_cb2d=native)
        B, T = int(fx["B"]), int(fx["T"])
        g_keys = [str(k) for k in fx["g_keys"]]
        for step in (1, 2):
            real = synth_mels(B, T, cfg.mel_channels, seed=40 + step)
            lens = synth_lengths(B, T, seed=40 + step, ragged=True)
            real = real.masked_fill((torch.arange(T)[None, :] >= lens[:, None]).unsqueeze(-1), 0.0)
            o = ts.step(real, lens, gan=True, use_fm=step == 2)
            pre = f"s{step}_"
            names = ["loss_d", "loss_g_total", "loss_recon_pre", "loss_recon_post", "loss_gan", "loss_fm"]
            got = np.array([float(o[n]) for n in names])
            ref = fx[pre + "losses"]
            r = {"losses": got.tolist(), "ref_losses": ref.tolist(),
                 "loss_rel_err": (np.abs(got - ref) / np.maximum(np.abs(ref), 1e-6)).tolist()}
            rp, rq = ts.last_recon
            r["recon_pre_max_abs_err"] = float((rp.cpu() - torch.from_numpy(fx[pre + "recon_pre"])).abs().max())
            r["recon_post_max_abs_err"] = float((rq.cpu() - torch.from_numpy(fx[pre + "recon_post"])).abs().max())
            gn = np.array([float(ts.g[k].grad.norm()) for k in g_keys])
            refn = fx[pre + "g_grad_norms"]
            has = refn > 1e-4 * refn.max()          # gradients that are float noise in the reference (e.g. d/dv of v/|v| for a
            rel = np.abs(gn[has] - refn[has]) / refn[has]      # 1-element v, pre.pw original1) carry no information
            r["g_grad_norm_rel_err_max"] = float(rel.max())
            r["g_grad_norm_rel_err_median"] = float(np.median(rel))
            r["g_grad_norm_worst"] = g_keys[int(np.flatnonzero(has)[int(rel.argmax())])]
            r["no_grad_params_zero"] = bool(all(float(ts.g[k].grad.abs().max()) == 0.0 for k, h in zip(g_keys, refn < 0) if h))
            r["noise_grad_abs_max"] = float(max(gn[~has & (refn >= 0)], default=0.0))
            cos = {}
            l2 = {}
            for name in fx.files:
                if name.startswith(pre + "gg:"):
                    k = name[len(pre) + 3:]
                    a = ts.g[k].grad.detach().cpu().double().reshape(-1)
                    b = torch.from_numpy(fx[name]).double().reshape(-1)
                    if float(b.norm()) <= 1e-4 * refn.max():
                        continue
                    cos[k] = float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))
                    l2[k] = float((a - b).norm() / b.norm())
            r["g_grad_cosine_min"] = min(cos.values())
            r["g_grad_cosine_argmin"] = min(cos, key=cos.get)
            r["g_grad_rel_l2_max"] = max(l2.values())
            r["g_grad_rel_l2_argmax"] = max(l2, key=l2.get)
            r["g_grad_rel_l2"] = l2
            ps = np.array([float(ts.g[k].detach().double().sum()) for k in g_keys])
            r["g_param_sum_abs_err_max"] = float(np.abs(ps - fx[pre + "g_param_sums"]).max())
            r["lecam"] = ts.lecam.ema.tolist()
            r["ref_lecam"] = fx[pre + "lecam"].tolist()
            r["u_err"] = float((ts.pd["convs.1.weight_u"].cpu() - torch.from_numpy(fx[pre + "d_u0"])).abs().max())
            d_now = {**{"pd:" + k: v for k, v in ts.pd.items()}, **{"mb:" + k: v for k, v in ts.mb.items()}}
            dps = np.array([float(d_now[str(k)].detach().double().sum()) for k in fx["d_keys"]])
            r["d_param_sum_abs_err_max"] = float(np.abs(dps - fx[pre + "d_param_sums"]).max())
            rep[f"{fixture}_{'native' if native else 'torch'}_cb2d_step{step}"] = r
    return rep


if __name__ == "__main__":
    main()
