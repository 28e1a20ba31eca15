"""transformer_tacotron2_b200 -- B200-native (sm_100a) Transformer-TTS hot path.

`TransformerTTS` mirrors the module API BASELINE.json's north_star defines (forward / inference,
same state_dict keys); the arithmetic is in libtts_b200.so (csrc/, C ABI in include/tts_b200.h).
"""
from .model import TTSConfig, TransformerTTS  # noqa: F401
from . import _lib  # noqa: F401

__all__ = ["TTSConfig", "TransformerTTS"]
