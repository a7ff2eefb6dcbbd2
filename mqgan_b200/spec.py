"""Architecture description of the PreEncoder hot path.

One place that knows the layer shapes and the state-dict key names, so the
nn.Module boundary (``preencoder.py``), the weight packer (``engine.py``) and
the synthetic-weight generator (``synth.py``) agree with the reference's
checkpoints key for key.

Reference: preencoder.py:304-361 (PreEncoder.__init__), :134-166 (UNetRefiner),
:205-268 (ConvBlock2D), attentions.py:476-523 (ResidualBlock1D),
:195-215 (CAM1D), :284-308 (SAM1D).  Two weight-norm flavours coexist
(SURVEY App. B4): ``parametrizations.weight.original0/1`` (g, v) for the
encoder / ConvBlock2D / refiner and legacy ``weight_g`` / ``weight_v`` for the
causal decoder convs.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

Shape = Tuple[int, ...]


@dataclass(frozen=True)
class PreEncoderConfig:
    mel_channels: int
    channels: Tuple[int, ...]
    kernel_sizes: Tuple[int, ...]
    fsq_levels: Tuple[int, ...] = (8, 8, 5, 5, 5)
    refiner_base_channels: int = 128
    refiner_depth: int = 3
    refiner_hidden_proj_divisor: int = 8

    def __post_init__(self):
        # The reference indexes kernel_sizes[i] for i < len(channels)-1 and the
        # *reversed* list for the decoder (preencoder.py:326,344-347), so the
        # shipped configs carry one more kernel size than blocks: encoder uses
        # [3,3,5] and the decoder [7,5,3] of kernel_sizes [3,3,5,7].
        if len(self.kernel_sizes) < len(self.channels) - 1:
            raise IndexError("kernel_sizes needs at least len(channels) - 1 entries")

    # ---- derived sizes -------------------------------------------------
    @property
    def c0(self) -> int:
        return self.channels[0]

    @property
    def latent_dim(self) -> int:
        return self.channels[-1]

    @property
    def quantizer_dim(self) -> int:
        return len(self.fsq_levels)

    @property
    def codebook_size(self) -> int:
        n = 1
        for lv in self.fsq_levels:
            n *= lv
        return n

    @property
    def refiner_hidden_channels(self) -> int:
        return self.mel_channels // self.refiner_hidden_proj_divisor

    @property
    def refiner_width(self) -> int:
        """F axis of the refiner image = mel + hidden channels (preencoder.py:358-359)."""
        return self.mel_channels + self.refiner_hidden_channels

    @property
    def refiner_channels(self) -> Tuple[int, ...]:
        return tuple(self.refiner_base_channels * (2 ** i) for i in range(self.refiner_depth + 1))

    @property
    def encoder_layers(self) -> List[Tuple[int, int, int]]:
        """(c_in, c_out, kernel) per encoder ResidualBlock1D."""
        return [(self.channels[i], self.channels[i + 1], self.kernel_sizes[i])
                for i in range(len(self.channels) - 1)]

    @property
    def decoder_layers(self) -> List[Tuple[int, int, int]]:
        rc = list(reversed(self.channels))
        rk = list(reversed(self.kernel_sizes))
        return [(rc[i], rc[i + 1], rk[i]) for i in range(len(rc) - 1)]

    @staticmethod
    def from_yaml_dict(cfg: dict) -> "PreEncoderConfig":
        """Build from a reference model_config*.yaml dict (same .get defaults as
        reencode_spectrograms_from_checkpoint.py:27-37)."""
        mp = cfg["model"]
        gp = mp["generator"]
        return PreEncoderConfig(
            mel_channels=int(mp["mel_channels"]),
            channels=tuple(gp["channels"]),
            kernel_sizes=tuple(gp["kernel_sizes"]),
            fsq_levels=tuple(gp["fsq_levels"]),
            refiner_base_channels=int(gp.get("refiner_base_channels", 128)),
            refiner_depth=int(gp.get("refiner_depth", 3)),
            refiner_hidden_proj_divisor=int(gp.get("refiner_hidden_proj_divisor", 8)),
        )


HIFISPEECH = PreEncoderConfig(128, (512, 512, 512, 768), (3, 3, 5, 7), (8, 5, 5, 5), 64, 3, 8)
HIFIMUSIC = PreEncoderConfig(160, (384, 384, 512, 512), (3, 3, 5, 7), (8, 5, 5, 5), 96, 3, 8)
# A small model with the same topology (channel change in the last encoder /
# first decoder block, 3-level refiner) for fast CPU-side tests.
TINY = PreEncoderConfig(32, (64, 64, 64, 128), (3, 3, 5, 7), (8, 5, 5, 5), 16, 3, 8)


def _wn_new(prefix: str, w_shape: Shape) -> List[Tuple[str, Shape]]:
    g_shape = (w_shape[0],) + (1,) * (len(w_shape) - 1)
    return [
        (prefix + ".bias", (w_shape[0],)),
        (prefix + ".parametrizations.weight.original0", g_shape),
        (prefix + ".parametrizations.weight.original1", w_shape),
    ]


def _wn_old(prefix: str, w_shape: Shape) -> List[Tuple[str, Shape]]:
    g_shape = (w_shape[0],) + (1,) * (len(w_shape) - 1)
    return [
        (prefix + ".bias", (w_shape[0],)),
        (prefix + ".weight_g", g_shape),
        (prefix + ".weight_v", w_shape),
    ]


def _plain(prefix: str, w_shape: Shape, bias: bool = True) -> List[Tuple[str, Shape]]:
    out = [(prefix + ".weight", w_shape)]
    if bias:
        out.append((prefix + ".bias", (w_shape[0],)))
    return out


def _convblock2d(prefix: str, c: int) -> List[Tuple[str, Shape]]:
    return (
        _wn_new(prefix + ".dw", (1, 1, 5, 5))
        + _wn_new(prefix + ".pw", (c, 1, 1, 1))
        + _plain(prefix + ".conv_out", (1, c, 1, 1))
    )


def _refiner_convblock(prefix: str, cin: int, cout: int) -> List[Tuple[str, Shape]]:
    return _wn_new(prefix + ".conv1", (cout, cin, 3, 3)) + _wn_new(prefix + ".conv2", (cout, cout, 3, 3))


def param_spec(cfg: PreEncoderConfig) -> List[Tuple[str, Shape]]:
    """Ordered (state-dict key, shape) list, in the reference's registration order."""
    s: List[Tuple[str, Shape]] = []
    s += _plain("proj", (cfg.c0, cfg.mel_channels))
    s += _convblock2d("pre", cfg.c0)
    for i, (cin, cout, k) in enumerate(cfg.encoder_layers):
        p = f"encoder_blocks.{i}"
        s += _wn_new(p + ".conv1", (cout, cin, k))
        s += _wn_new(p + ".conv2", (cout, cout, k))
        r = cout // 8
        s += _plain(p + ".cbam.channel_attention.mlp.0", (r, cout))
        s += _plain(p + ".cbam.channel_attention.mlp.2", (cout, r))
        s += _plain(p + ".cbam.spatial_attention.conv", (1, 2, 7), bias=False)
        s += [(p + ".relu.beta", ()), (p + ".relu.gamma", ())]
        if cin != cout:
            s += _plain(p + ".residual", (cout, cin, 1))
    s += _plain("q_in_proj", (cfg.quantizer_dim, cfg.latent_dim))
    s += _plain("q_out_proj", (cfg.latent_dim, cfg.quantizer_dim))
    for i, (cin, cout, k) in enumerate(cfg.decoder_layers):
        p = f"decoder_blocks.{i}"
        s += _wn_old(p + ".conv1", (cout, cin, k))
        s += _wn_old(p + ".conv2", (cout, cout, k))
        s += [(p + ".relu.beta", ()), (p + ".relu.gamma", ())]
        if cin != cout:
            s += _plain(p + ".residual", (cout, cin, 1))
    s += _convblock2d("post", cfg.c0)
    s += _plain("out_proj", (cfg.mel_channels, cfg.c0))
    s += _plain("hidden_proj", (cfg.refiner_hidden_channels, cfg.c0))
    chs = cfg.refiner_channels
    d = cfg.refiner_depth
    s += _refiner_convblock("refiner.pre", 1, chs[0])
    for i in range(d):
        s += _refiner_convblock(f"refiner.downs.{i}.conv", chs[i], chs[i + 1])
    s += _refiner_convblock("refiner.mid", chs[-1], chs[-1])
    for i in range(d):
        cin = chs[d - i] + chs[d - i - 1]
        s += _refiner_convblock(f"refiner.ups.{i}.conv", cin, chs[d - i - 1])
    s += _wn_new("refiner.post", (1, chs[0], 3, 3))
    s += _plain("refiner.reproj", (cfg.mel_channels, cfg.refiner_width), bias=False)
    return s


def flops_per_frame(cfg: PreEncoderConfig) -> Dict[str, float]:
    """Algorithmic FLOPs (2 per MAC of every conv / linear) per mel frame at
    T % 2**depth == 0 — the figure SURVEY §8(d) quotes (738.78 M for hifispeech)."""
    enc = 2.0 * cfg.mel_channels * cfg.c0
    pp = 2.0 * cfg.c0 * (25 + 2 * cfg.c0)  # ConvBlock2D: dw 25 MAC + pw C + conv_out C per pixel
    enc += pp
    for cin, cout, k in cfg.encoder_layers:
        enc += 2.0 * k * (cin * cout + cout * cout)
        r = cout // 8
        enc += 0.0 * r  # CAM MLP acts on a (B,C) vector: not per frame
        enc += 2.0 * 14  # SAM conv 2->1, k=7
        if cin != cout:
            enc += 2.0 * cin * cout
    enc += 2.0 * cfg.latent_dim * cfg.quantizer_dim
    dec = 2.0 * cfg.latent_dim * cfg.quantizer_dim
    for cin, cout, k in cfg.decoder_layers:
        dec += 2.0 * k * (cin * cout + cout * cout)
        if cin != cout:
            dec += 2.0 * cin * cout
    dec += pp
    dec += 2.0 * cfg.c0 * (cfg.mel_channels + cfg.refiner_hidden_channels)
    F = cfg.refiner_width
    chs = cfg.refiner_channels
    d = cfg.refiner_depth
    ref = 0.0

    def cb(cin, cout, s):
        return 2.0 * 9 * (cin * cout + cout * cout) * F / s

    ref += cb(1, chs[0], 1)
    for i in range(d):
        ref += cb(chs[i], chs[i + 1], 2 ** (i + 1))
    ref += cb(chs[-1], chs[-1], 2 ** d)
    for i in range(d):
        ref += cb(chs[d - i] + chs[d - i - 1], chs[d - i - 1], 2 ** (d - i - 1))
    ref += 2.0 * 9 * chs[0] * F
    ref += 2.0 * F * cfg.mel_channels
    return {"encode": enc, "decode_1d": dec, "refiner": ref, "total": enc + dec + ref}


# ----------------------------------------------------------------------------
# training step: discriminators and the reference's training hyper-parameters
# ----------------------------------------------------------------------------
@dataclass(frozen=True)
class PatchDiscConfig:
    """MelSpectrogramPatchDiscriminator2D (discriminators.py:70-190): ``len(hidden_channels)`` strided,
    spectrally-normalised Conv2d layers over the (mel, time) plane plus a stride-1 logits conv, LeakyReLU(0.2)
    after every conv, a masked squeeze-excite (reduction 8) in front of the last one."""
    mel_channels: int
    hidden_channels: Tuple[int, ...]
    kernels: Tuple[Tuple[int, int], ...]     # (kh, kw) per conv, len(hidden_channels) + 1
    strides: Tuple[Tuple[int, int], ...]     # (sh, sw) per conv; the logits conv always runs at stride 1

    def __post_init__(self):
        if len(self.kernels) != len(self.hidden_channels) + 1 or len(self.strides) != len(self.kernels):
            raise ValueError("kernel_sizes / strides must have len(hidden_channels) + 1 entries")

    @staticmethod
    def from_patch_yaml(mel_channels: int, d: dict) -> "PatchDiscConfig":
        """``discriminator_patch`` section of a reference model_config*.yaml (train.py:290-296)."""
        return PatchDiscConfig(int(mel_channels), tuple(d["hidden_channels"]),
                               tuple((int(k), int(k)) for k in d["kernel_sizes"]),
                               tuple((int(s[0]), int(s[1])) for s in d["strides"]))

    def conv_shapes(self) -> List[Shape]:
        cin, out = 1, []
        for c, (kh, kw) in zip(tuple(self.hidden_channels) + (1,), self.kernels):
            out.append((c, cin, kh, kw))
            cin = c
        return out

    def layer_stride(self, i: int) -> Tuple[int, int]:
        return self.strides[i] if i < len(self.kernels) - 1 else (1, 1)

    @property
    def feature_layers(self) -> Tuple[bool, ...]:
        """Which conv outputs feature matching uses (discriminators.py:108-112)."""
        n = len(self.kernels)
        return tuple(not (i in (0, 1) or i == n - 1) for i in range(n))


@dataclass(frozen=True)
class MultiBinConfig:
    """MultiBinDiscriminator (discriminators.py:260-312): ``n_bins`` equal mel bands, one patch
    discriminator each with (3, k) kernels, no stride on the first ``n_no_strides`` layers, then (1, 2)."""
    mel_channels: int
    n_bins: int
    hidden_channels: Tuple[int, ...]
    kernel_sizes: Tuple[int, ...]
    n_no_strides: int = 2

    def __post_init__(self):
        # discriminators.py:262-266
        if self.mel_channels % self.n_bins:
            raise ValueError("mel_channels must divide n_bins")
        for h in self.hidden_channels:
            if h % self.n_bins:
                raise ValueError(f"hidden size {h} must divide n_bins")

    @staticmethod
    def from_yaml(mel_channels: int, d: dict) -> "MultiBinConfig":
        return MultiBinConfig(int(mel_channels), int(d["n_bins"]), tuple(d["hidden_channels"]),
                              tuple(int(k) for k in d["kernel_sizes"]), int(d.get("n_no_strides", 2)))

    @property
    def bin_config(self) -> PatchDiscConfig:
        n = len(self.kernel_sizes)
        return PatchDiscConfig(self.mel_channels // self.n_bins, tuple(self.hidden_channels),
                               tuple((3, int(k)) for k in self.kernel_sizes),
                               tuple((1, 1) if i < self.n_no_strides else (1, 2) for i in range(n)))


def patch_disc_param_spec(dc: PatchDiscConfig, prefix: str = "") -> List[Tuple[str, Shape]]:
    """(state-dict key, shape) of one patch discriminator: legacy ``spectral_norm`` keeps ``weight_orig``
    as the parameter and ``weight_u`` / ``weight_v`` as buffers."""
    s: List[Tuple[str, Shape]] = []
    for i, w in enumerate(dc.conv_shapes()):
        p = f"{prefix}convs.{i}"
        s += [(p + ".bias", (w[0],)), (p + ".weight_orig", w), (p + ".weight_u", (w[0],)),
              (p + ".weight_v", (w[1] * w[2] * w[3],))]
    c = dc.hidden_channels[-1]
    r = max(1, c // 8)
    s += [(prefix + "se_block.fc1.weight", (r, c)), (prefix + "se_block.fc1.bias", (r,)),
          (prefix + "se_block.fc2.weight", (c, r)), (prefix + "se_block.fc2.bias", (c,))]
    return s


def multibin_param_spec(mc: MultiBinConfig) -> List[Tuple[str, Shape]]:
    s: List[Tuple[str, Shape]] = []
    for b in range(mc.n_bins):
        s += patch_disc_param_spec(mc.bin_config, f"discriminators.{b}.")
    return s


def is_disc_buffer(key: str) -> bool:
    return key.endswith(".weight_u") or key.endswith(".weight_v")


# training hyper-parameters of configs/model_config_hifispeech.yaml:31-48 (``training`` section)
TRAIN_DEFAULTS = {
    "lr": 1e-4, "beta1": 0.9, "beta2": 0.999, "lr_d_factor": 1.15, "d_beta1": 0.5, "d_beta2": 0.999,
    "warmup_steps": 1000, "discriminator_train_start_epoch": 8, "use_fm_loss": False,
    "loss_weights": {"fm_lambda": 0.25, "Gloss_lambda": 15.0, "recon_lambda": 15.0},
}
HIFISPEECH_PATCH_D = PatchDiscConfig(128, (256, 256, 384, 512, 512), ((5, 5), (5, 5), (5, 5), (3, 3), (3, 3), (3, 3)),
                                     ((1, 2), (2, 2), (2, 2), (2, 1), (2, 1), (2, 1)))
HIFISPEECH_MULTIBIN_D = MultiBinConfig(128, 8, (128, 128, 256, 256, 384), (7, 5, 3, 3, 3, 3), 2)
# configs/model_config_hifimusic.yaml:22-31
HIFIMUSIC_PATCH_D = PatchDiscConfig(160, (384, 384, 512, 512, 512), ((7, 7), (7, 7), (5, 5), (5, 5), (3, 3), (3, 3)),
                                    ((1, 2), (2, 2), (2, 2), (2, 2), (2, 2), (2, 2)))
HIFIMUSIC_MULTIBIN_D = MultiBinConfig(160, 8, (128, 256, 256, 256, 256), (7, 5, 5, 3, 3, 3), 2)
# small discriminators of the same topology for the TINY generator
TINY_PATCH_D = PatchDiscConfig(32, (16, 16, 32), ((5, 5), (5, 5), (3, 3), (3, 3)), ((1, 2), (2, 2), (2, 1), (1, 1)))
TINY_MULTIBIN_D = MultiBinConfig(32, 2, (16, 16, 32), (7, 5, 3, 3), 2)
TINY_TRAIN = dict(TRAIN_DEFAULTS, warmup_steps=4)
# A second small training topology with hifimusic's block pattern (channel change in the middle encoder / decoder
# block) and awkward widths: refiner channels 24/48/96/192 (not multiples of 64; 24 has no fused bias-gradient path),
# refiner image width 54 (padded to 56 channels for reproj), three 16-bin bands.
TINY_M = PreEncoderConfig(48, (48, 48, 64, 64), (3, 3, 5, 7), (8, 5, 5, 5), 24, 3, 8)
TINY_M_PATCH_D = PatchDiscConfig(48, (16, 24, 32), ((5, 5), (5, 5), (3, 3), (3, 3)), ((1, 2), (2, 2), (2, 1), (1, 1)))
TINY_M_MULTIBIN_D = MultiBinConfig(48, 3, (24, 24, 48), (7, 5, 3, 3), 2)
