"""ORACLE -- test infrastructure only (see transformer_tts.py header).  Never imported by the product
package `transformer_tacotron2_b200`; only tests/, __graft_entry__.smoke() and bench.py's CPU legs use it."""
from .transformer_tts import TTSConfig, TransformerTTS, tts_loss, sinusoid_table, length_mask  # noqa: F401
from . import philox, synthetic  # noqa: F401
