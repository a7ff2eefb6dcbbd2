"""Drop-in ``LSGANLoss`` / ``MaskedMelLoss`` (reference: losses.py:5-182) - same constructor arguments, methods, buffers
(``ema_real`` / ``ema_fake``) and arithmetic; plain torch ops, any device."""
from __future__ import annotations

import torch
import torch.nn as nn


class LSGANLoss(nn.Module):
    def __init__(self, real_label=1.0, fake_label=0.0, decay=0.99, use_lecam=True):
        super().__init__()
        self.real_label, self.fake_label, self.decay, self.use_lecam = real_label, fake_label, decay, use_lecam
        self.register_buffer("ema_real", torch.tensor(0.0))
        self.register_buffer("ema_fake", torch.tensor(0.0))
        self.ema_initialized = False

    def _masked_mse(self, pred, target, mask=None):
        """losses.py:21-35; mask True = valid.  No host read: an all-False mask yields 0 as the reference's branch does."""
        err = (pred - target) ** 2
        if mask is None:
            return err.mean()
        m = mask.float()
        valid = m.sum()
        return torch.where(valid > 0, (err * m).sum() / valid.clamp(min=1), err.new_zeros(()))

    @staticmethod
    def _masked_mean(x, mask):
        if mask is None:
            return x.mean()
        m = mask.float()
        return (x * m).sum() / m.sum().clamp(min=1)

    def update_ema(self, real_out, fake_out, real_mask=None, fake_mask=None):
        real_mean, fake_mean = self._masked_mean(real_out, real_mask).detach(), self._masked_mean(fake_out, fake_mask).detach()
        if not self.ema_initialized:                                                 # :50-53
            self.ema_real.copy_(real_mean)
            self.ema_fake.copy_(fake_mean)
            self.ema_initialized = True
        else:
            self.ema_real.mul_(self.decay).add_((1 - self.decay) * real_mean)
            self.ema_fake.mul_(self.decay).add_((1 - self.decay) * fake_mean)

    def lecam_loss(self, real_out, fake_out, real_mask=None, fake_mask=None):
        ema_r, ema_f = self.ema_real.detach().to(real_out.device), self.ema_fake.detach().to(real_out.device)
        term_r = self._masked_mean((real_out - ema_f).clamp(min=0) ** 2, real_mask)
        term_f = self._masked_mean((ema_r - fake_out).clamp(min=0) ** 2, fake_mask)
        return term_r + term_f

    def discriminator_loss(self, real_output, fake_output, real_mask=None, fake_mask=None):
        loss = 0.5 * (self._masked_mse(real_output, torch.full_like(real_output, self.real_label), real_mask)
                      + self._masked_mse(fake_output, torch.full_like(fake_output, self.fake_label), fake_mask))
        if self.use_lecam:                                                           # the EMA moves first (:96-99)
            self.update_ema(real_output, fake_output, real_mask, fake_mask)
            loss = loss + self.lecam_loss(real_output, fake_output, real_mask, fake_mask)
        return loss

    def generator_loss(self, fake_output, fake_mask=None):
        return self._masked_mse(fake_output, torch.full_like(fake_output, self.real_label), fake_mask)


class MaskedMelLoss(nn.Module):
    """Masked Charbonnier or MSE over mel bins, averaged per frequency group first (losses.py:126-182)."""

    def __init__(self, loss_type: str = "charbonnier", group_size: int = 1, eps: float = 1e-6):
        super().__init__()
        assert loss_type in {"charbonnier", "mse"}
        self.loss_type, self.group_size, self.eps = loss_type, group_size, eps

    def forward(self, x, y, lengths):
        assert x.shape == y.shape, "x and y must have the same shape"
        B, T, C = x.shape
        g = self.group_size
        assert C % g == 0, "C (n_mels) must be divisible by group_size"
        G = C // g
        mask = (torch.arange(T, device=x.device)[None, :] >= lengths.to(x.device)[:, None])[:, :, None].expand(B, T, C)
        mask = mask.reshape(B, T, G, g)
        diff = (x - y).reshape(B, T, G, g)
        per = torch.sqrt(diff.pow(2) + self.eps ** 2) if self.loss_type == "charbonnier" else diff.pow(2)
        per = per.masked_fill(mask, 0.0)
        return (per.sum(dim=[0, 1, 3]) / ((~mask).float().sum(dim=[0, 1, 3]) + 1e-12)).mean()
