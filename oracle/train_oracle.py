"""CPU oracle for the TRAINING step (BASELINE configs[4]).  TEST INFRASTRUCTURE ONLY.

Only ``tests/`` and the ``cpu_baseline`` / ``--impl reference`` legs of the benchmark tools may import
this module.  Nothing under ``mqgan_b200/`` imports it.

Functional restatement (state-dicts in, losses / gradients / updated state out; plain torch on the CPU,
autograd for the derivatives, no code from the reference) of one iteration of the reference's loop:

    train.py:521-529        _train_epoch body: G forward, D step, G step
    preencoder.py:363-418   PreEncoder.forward (training; refiner input detached :411-413)
    quantizer.py:128-140    FSQ.quantize with the straight-through round (noise_dropout = 0)
    discriminators.py:70-257   MelSpectrogramPatchDiscriminator2D (+ ChannelSELayerMasked :10-67)
    discriminators.py:260-312  MultiBinDiscriminator
    torch.nn.utils.spectral_norm   one power iteration per training-mode forward, none in eval
    losses.py:5-121         LSGANLoss (masked MSE, LeCam with EMA)
    losses.py:126-182       MaskedMelLoss("mse", group_size)
    train.py:38-45          masked_mae (feature matching)
    train.py:380-412        _train_discriminator (bin 0's mask for every bin, App. B13)
    train.py:414-501        _train_generator
    train.py:314-329        Adam(G), Adam(D, lr * lr_d_factor, d betas), LambdaLR warm-up

Dropout is taken as 0 EVERYWHERE, including the 0.1 that DownBlock / UpBlock / ConvBlock2D hard-wire
regardless of the ``dropout`` argument (preencoder.py:109, 121, 233): the reference draws it from torch's
global RNG, which no other implementation can reproduce.  Everything else follows the reference, including
its quirks: the discriminators see ``recon_post`` only, ``_train_generator`` puts them in eval mode and never back (so from the second
iteration on the power iteration no longer runs - train.py:417-418, :504-506 resets it per epoch only),
and the LeCam EMA is updated before it is used (losses.py:96-99).

Parity pin: ``oracle/make_golden_train.py`` drives the UNMODIFIED reference (its own ``PreEncoder``,
discriminators, losses and the two ``Trainer`` methods, called on a stand-in ``self``) in the build
container and commits losses, gradient norms and updated-parameter checksums to
``tests/golden/train_tiny.npz``; ``tests/test_train_oracle.py`` re-checks this module against them.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import preencoder_oracle as O

Tensor = torch.Tensor


# ----------------------------------------------------------------------------
# generator, training mode
# ----------------------------------------------------------------------------
def effective_weights_grad(params: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """As preencoder_oracle.effective_weights but differentiable w.r.t. g and v."""
    out: Dict[str, Tensor] = {}
    for k, t in params.items():
        if k.endswith(".parametrizations.weight.original1"):
            base = k[: -len(".parametrizations.weight.original1")]
            out[base + ".weight"] = torch._weight_norm(t, params[base + ".parametrizations.weight.original0"], 0)
        elif k.endswith(".weight_v"):
            base = k[: -len(".weight_v")]
            out[base + ".weight"] = torch._weight_norm(t, params[base + ".weight_g"], 0)
        elif k.endswith("original0") or k.endswith(".weight_g"):
            continue
        else:
            out[k] = t
    return out


def fsq_quantize_ste(z: Tensor, levels: Sequence[int]) -> Tensor:
    """quantizer.py:128-140 in training mode with noise_dropout = 0: round with a straight-through
    gradient (round_ste), divided by the half width."""
    lv, basis, half_l, offset, shift, half_w = O.fsq_constants(levels, z.dtype)
    bounded = (z + shift).tanh() * half_l - offset
    q = bounded + (bounded.round() - bounded).detach()
    return q / half_w


def generator_forward(params: Dict[str, Tensor], cfg, mel: Tensor, lengths: Tensor) -> Tuple[Tensor, Tensor]:
    """PreEncoder.forward (preencoder.py:363-418) -> (x_recon, x_post)."""
    w = effective_weights_grad(params)
    T = mel.shape[1]
    mask_bt = O.sequence_mask(T, lengths)
    mask = mask_bt.unsqueeze(1)
    x = F.linear(mel, w["proj.weight"], w["proj.bias"]).permute(0, 2, 1)
    x = O.convblock2d(x, mask, w, "pre")
    for i, (_, _, k) in enumerate(cfg.encoder_layers):
        x = O.residual_block(x, mask, w, f"encoder_blocks.{i}", k, causal=False)
    z = F.linear(x.permute(0, 2, 1), w["q_in_proj.weight"], w["q_in_proj.bias"])
    codes = fsq_quantize_ste(z, cfg.fsq_levels)
    dec = F.linear(codes, w["q_out_proj.weight"], w["q_out_proj.bias"]).permute(0, 2, 1)
    for i, (_, _, k) in enumerate(cfg.decoder_layers):
        dec = O.residual_block(dec, mask, w, f"decoder_blocks.{i}", k, causal=True)
    xr = O.convblock2d(dec, mask, w, "post").permute(0, 2, 1)
    x_recon = F.linear(xr, w["out_proj.weight"], w["out_proj.bias"])
    hid = F.linear(dec.permute(0, 2, 1), w["hidden_proj.weight"], w["hidden_proj.bias"])
    r_in = torch.cat([x_recon, hid], dim=2).unsqueeze(1).detach()           # :411-413
    res = O.refiner(r_in, mask_bt, w, cfg.refiner_depth)
    return x_recon, x_recon + res


# ----------------------------------------------------------------------------
# discriminators
# ----------------------------------------------------------------------------
def spectral_weight(sd: Dict[str, Tensor], prefix: str, training: bool) -> Tensor:
    """torch.nn.utils.spectral_norm (legacy hook, n_power_iterations = 1, eps = 1e-12, dim 0): in
    training mode u, v advance one power iteration IN PLACE (buffers), then w = w_orig / (u . W v)."""
    w_orig = sd[prefix + ".weight_orig"]
    u, v = sd[prefix + ".weight_u"], sd[prefix + ".weight_v"]
    wm = w_orig.reshape(w_orig.shape[0], -1)
    if training:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=1e-12))
            u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=1e-12))
        u, v = u.clone(), v.clone()
    sigma = torch.dot(u, torch.mv(wm, v))
    return w_orig / sigma


def patch_discriminator(sd: Dict[str, Tensor], dcfg: dict, x: Tensor, lengths: Tensor, training: bool,
                        prefix: str = "") -> Tuple[Tensor, Tensor, List[Tuple[Tensor, Tensor]]]:
    """MelSpectrogramPatchDiscriminator2D.forward (discriminators.py:208-257) ->
    (logits (B,1,H,W), valid-patch mask, [(feature, padded mask)] for the layers :108-112 selects).
    dcfg: {"kernels": [(kh, kw)], "strides": [(sh, sw)]} per conv, the last one being the logits conv."""
    B, T, Fm = x.shape
    kernels, strides = dcfg["kernels"], dcfg["strides"]
    n = len(kernels)
    pad_mask = O.sequence_mask(T, lengths)[:, None, None, :].expand(-1, 1, Fm, -1)      # :200-205
    out = x.transpose(1, 2).unsqueeze(1)                                                 # (B,1,F,T) :226
    feats: List[Tuple[Tensor, Tensor]] = []
    ret = [True] * n
    ret[0] = ret[1] = ret[-1] = False
    for i in range(n):
        if i == n - 1:                                                                   # SE before the logits conv :231-232
            valid = ~pad_mask
            denom = valid.sum(dim=(2, 3)).clamp(min=1)
            squeeze = (out * valid).reshape(B, out.shape[1], -1).sum(dim=2) / denom
            h = F.relu(F.linear(squeeze, sd[prefix + "se_block.fc1.weight"], sd[prefix + "se_block.fc1.bias"]))
            ex = torch.sigmoid(F.linear(h, sd[prefix + "se_block.fc2.weight"], sd[prefix + "se_block.fc2.bias"]))
            out = out * ex.reshape(B, -1, 1, 1)
        kh, kw = kernels[i]
        sh, sw = strides[i] if i < n - 1 else (1, 1)
        w = spectral_weight(sd, f"{prefix}convs.{i}", training)
        out = F.leaky_relu(F.conv2d(out, w, sd[f"{prefix}convs.{i}.bias"], stride=(sh, sw),
                                    padding=((kh - 1) // 2, (kw - 1) // 2)), 0.2)        # :234
        if sh > 1 or sw > 1:                                                             # :237-244
            pad_mask = F.max_pool2d(pad_mask.float(), kernel_size=(sh, sw), stride=(sh, sw), ceil_mode=True).bool()
        out = out.masked_fill(pad_mask, 0.0)                                             # :247
        if ret[i]:
            feats.append((out, pad_mask))
    return out, ~pad_mask, feats


def multibin_discriminator(sd: Dict[str, Tensor], dcfg: dict, x: Tensor, lengths: Tensor, training: bool):
    """MultiBinDiscriminator.forward (discriminators.py:292-312): equal mel bands, one patch discriminator each."""
    n_bins = dcfg["n_bins"]
    outs, masks, feats = [], [], []
    for b, sub in enumerate(torch.split(x, x.size(-1) // n_bins, dim=-1)):
        o, m, f = patch_discriminator(sd, dcfg, sub, lengths, training, prefix=f"discriminators.{b}.")
        outs.append(o); masks.append(m); feats.append(f)
    return outs, masks, feats


def patch_cfg(kernel_sizes, strides) -> dict:
    """discriminators.py:114-143: square kernels, per-layer (h, w) strides."""
    return {"kernels": [(k, k) for k in kernel_sizes], "strides": [tuple(s) for s in strides]}


def multibin_cfg(kernel_sizes, n_bins, n_no_strides) -> dict:
    """discriminators.py:276-289: (3, k) kernels; stride (1,1) for the first n_no_strides layers, then (1,2)."""
    n = len(kernel_sizes)
    return {"kernels": [(3, k) for k in kernel_sizes],
            "strides": [(1, 1) if i < n_no_strides else (1, 2) for i in range(n)], "n_bins": n_bins}


# ----------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------
def masked_mse(pred: Tensor, target: float, mask: Tensor) -> Tensor:
    """losses.py:21-35 (mask True = valid)."""
    m = mask.to(pred.dtype)
    valid = m.sum()
    if valid.item() > 0:
        return (((pred - target) ** 2) * m).sum() / valid
    return pred.new_zeros(())


class LeCamState:
    """EMA buffers of LSGANLoss (losses.py:17-20, 37-58)."""

    def __init__(self, decay: float = 0.99):
        self.decay = decay
        self.ema_real = torch.tensor(0.0)
        self.ema_fake = torch.tensor(0.0)
        self.initialized = False

    def update(self, real_mean: Tensor, fake_mean: Tensor):
        if not self.initialized:
            self.ema_real, self.ema_fake, self.initialized = real_mean.detach().clone(), fake_mean.detach().clone(), True
        else:
            self.ema_real = self.ema_real * self.decay + (1 - self.decay) * real_mean.detach()
            self.ema_fake = self.ema_fake * self.decay + (1 - self.decay) * fake_mean.detach()


def lsgan_d_loss(state: LeCamState, real: Tensor, fake: Tensor, real_mask: Tensor, fake_mask: Tensor) -> Tensor:
    """LSGANLoss.discriminator_loss (losses.py:81-107): the EMA moves first, then LeCam uses it."""
    loss = 0.5 * (masked_mse(real, 1.0, real_mask) + masked_mse(fake, 0.0, fake_mask))
    rm, fm = real_mask.to(real.dtype), fake_mask.to(fake.dtype)
    state.update((real * rm).sum() / rm.sum().clamp(min=1), (fake * fm).sum() / fm.sum().clamp(min=1))
    term_r = (((real - state.ema_fake).clamp(min=0) * rm) ** 2).sum() / rm.sum().clamp(min=1)
    term_f = (((state.ema_real - fake).clamp(min=0) * fm) ** 2).sum() / fm.sum().clamp(min=1)
    return loss + term_r + term_f


def masked_mel_loss(x: Tensor, y: Tensor, lengths: Tensor, group_size: int) -> Tensor:
    """MaskedMelLoss("mse", group_size).forward (losses.py:148-182)."""
    B, T, C = x.shape
    G = C // group_size
    pad = (torch.arange(T)[None, :] >= lengths[:, None])[:, :, None].expand(B, T, C).reshape(B, T, G, group_size)
    per = ((x - y) ** 2).reshape(B, T, G, group_size).masked_fill(pad, 0.0)
    group_sum = per.sum(dim=[0, 1, 3])
    count = (~pad).to(x.dtype).sum(dim=[0, 1, 3])
    return (group_sum / (count + 1e-12)).mean()


def masked_mae(pred: Tensor, target: Tensor, mask: Tensor, eps: float = 1e-8) -> Tensor:
    """train.py:38-45 (mask True = padded)."""
    mask = mask.expand_as(pred)
    diff = (pred - target).abs().masked_fill(mask, 0.0)
    return diff.sum() / ((~mask).sum() + eps)


# ----------------------------------------------------------------------------
# one iteration
# ----------------------------------------------------------------------------
class TrainState:
    """Parameters, buffers and optimiser state of one replica (all CPU tensors)."""

    def __init__(self, cfg, g_sd, pd_sd, mb_sd, pd_cfg, mb_cfg, tcfg: dict, dtype=torch.float32):
        self.cfg, self.pd_cfg, self.mb_cfg, self.tcfg = cfg, pd_cfg, mb_cfg, tcfg
        self.g = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in g_sd.items()}
        self.pd = {k: v.detach().clone().to(dtype) for k, v in pd_sd.items()}
        self.mb = {k: v.detach().clone().to(dtype) for k, v in mb_sd.items()}
        for sd in (self.pd, self.mb):
            for k, v in sd.items():
                if not (k.endswith("weight_u") or k.endswith("weight_v")):
                    v.requires_grad_(True)
        self.d_training = True               # the reference flips D to eval in the first G step and never back
        self.lecam = LeCamState()
        t = tcfg
        self.opt_g = torch.optim.Adam(list(self.g.values()), lr=t["lr"], betas=(t["beta1"], t["beta2"]))
        self.opt_d = torch.optim.Adam(self.d_params(), lr=t["lr"] * t["lr_d_factor"], betas=(t["d_beta1"], t["d_beta2"]))
        self.sched_g = torch.optim.lr_scheduler.LambdaLR(
            self.opt_g, lambda step: min((step + 1) / t["warmup_steps"], 1.0))

    def d_params(self) -> List[Tensor]:
        return [v for sd in (self.pd, self.mb) for v in sd.values() if v.requires_grad]


def train_iteration(st: TrainState, real: Tensor, lengths: Tensor, gan: bool = True, use_fm: bool = False) -> dict:
    """train.py:521-529 for one batch.  Returns the losses the reference logs; gradients stay in .grad
    (clipped, as the reference leaves them) and the parameters are updated in place."""
    t = st.tcfg
    clip = t.get("clip_grad_norm", 1.0)
    lw = t["loss_weights"]
    recon_pre, recon_post = generator_forward(st.g, st.cfg, real, lengths)
    out = {"loss_d": 0.0}
    if gan:                                                                       # _train_discriminator :380-412
        st.opt_d.zero_grad()
        rl, rm, _ = patch_discriminator(st.pd, st.pd_cfg, real, lengths, st.d_training)
        fl, fm, _ = patch_discriminator(st.pd, st.pd_cfg, recon_post.detach(), lengths, st.d_training)
        loss_d1 = lsgan_d_loss(st.lecam, rl, fl, rm, fm)
        rl2, rm2, _ = multibin_discriminator(st.mb, st.mb_cfg, real, lengths, st.d_training)
        fl2, fm2, _ = multibin_discriminator(st.mb, st.mb_cfg, recon_post.detach(), lengths, st.d_training)
        loss_mbd = sum(lsgan_d_loss(st.lecam, r, f, rm2[0], fm2[0]) for r, f in zip(rl2, fl2)) / len(rl2)
        loss_d = loss_d1 + loss_mbd
        loss_d.backward()
        if clip:
            torch.nn.utils.clip_grad_norm_(st.d_params(), clip)
        st.opt_d.step()
        out["loss_d"] = float(loss_d)
    # _train_generator :414-501
    st.opt_g.zero_grad()
    st.d_training = False                                                         # :417-418
    mel_all = lambda a: masked_mel_loss(a, real, lengths, 1)
    mel_grp = lambda a: masked_mel_loss(a, real, lengths, 16)
    loss_recon_pre = mel_all(recon_pre) + 0.25 * mel_grp(recon_pre)
    loss_recon_post = mel_all(recon_post) + 0.25 * mel_grp(recon_post)
    loss_gan = real.new_zeros(())
    loss_fm = real.new_zeros(())
    gl_lambda = fm_lambda = 0.0
    if gan:
        gl, gm, gf = patch_discriminator(st.pd, st.pd_cfg, recon_post, lengths, False)
        gl2, gm2, gf2 = multibin_discriminator(st.mb, st.mb_cfg, recon_post, lengths, False)
        loss_gan = 0.5 * (masked_mse(gl, 1.0, gm) + sum(masked_mse(g, 1.0, gm2[0]) for g in gl2) / len(gl2))
        gl_lambda, fm_lambda = lw["Gloss_lambda"], lw["fm_lambda"]
        if use_fm:                                                                # :454-476
            with torch.no_grad():
                _, _, rf = patch_discriminator(st.pd, st.pd_cfg, real, lengths, False)
                _, _, rf2 = multibin_discriminator(st.mb, st.mb_cfg, real, lengths, False)
            fm_d1 = sum(masked_mae(ff, r, m) for (r, m), (ff, _) in zip(rf, gf)) / max(len(rf), 1)
            fm_mbd = real.new_zeros(())
            for rfe, gfe in zip(rf2, gf2):                                        # running division, as :466-472 does it
                for (r, m), (ff, _) in zip(rfe, gfe):
                    fm_mbd = fm_mbd + masked_mae(ff, r, m)
                if len(rfe) > 0:
                    fm_mbd = fm_mbd / len(rfe)
            fm_mbd = fm_mbd / max(len(gf2), 1)
            loss_fm = 0.5 * (fm_d1 + fm_mbd)
    total = (loss_recon_pre * lw.get("recon_lambda_pre", 1.0) + loss_recon_post * lw.get("recon_lambda_post", 2.0)
             + loss_gan * gl_lambda + loss_fm * fm_lambda)
    total.backward()
    if clip:
        torch.nn.utils.clip_grad_norm_(list(st.g.values()), clip)
    st.opt_g.step()
    st.sched_g.step()
    out.update(loss_g_total=float(total), loss_recon_pre=float(loss_recon_pre), loss_recon_post=float(loss_recon_post),
               loss_gan=float(loss_gan), loss_fm=float(loss_fm))
    out["recon_pre"], out["recon_post"] = recon_pre.detach(), recon_post.detach()
    return out
