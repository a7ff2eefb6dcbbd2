"""mqgan_b200 — B200-native (sm_100a) PreEncoder re-encode path of ZDisket/MQGAN.

Public surface mirrors the reference: ``PreEncoder``, ``get_pre_encoder``,
``sequence_mask``, ``strip_weight_norm`` (preencoder.py) and
``ScriptedPreEncoder`` (scripted_preencoder.py).
"""
from .spec import PreEncoderConfig, HIFISPEECH, HIFIMUSIC, TINY  # noqa: F401

__all__ = ["PreEncoderConfig", "HIFISPEECH", "HIFIMUSIC", "TINY"]
__version__ = "0.1.0"
