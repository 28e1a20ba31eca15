"""ORACLE (test infrastructure) -- canonical synthetic weights and inputs.

There are no datasets or checkpoints (no network; the reference ships none), so every run uses
the fixed random-init model and seeded synthetic inputs defined here.

Weights (P17): torch.manual_seed(1234); default nn inits; BatchNorm running stats randomised
(mean ~ N(0, 0.1), var ~ U[0.5, 1.5]) and BN / LayerNorm affine randomised so BN folding and the
LN affine are exercised; alpha_enc = 0.5, alpha_dec = 0.25;
every floating tensor except BN statistics / BN affine / LayerNorm affine / biases is then
ROUNDED TO A bf16-REPRESENTABLE VALUE (kept in fp32), so the bf16 weights the B200 path stores
are exactly the oracle's weights and parity measures arithmetic, not weight quantisation.
Finally the stop head is planted: stop_linear.bias = `stop_bias` (-8.0 => never fires: throughput
runs decode all max_len frames; about -0.45 => utterances stop at scattered frames: parity runs).
"""
from __future__ import annotations

import hashlib
from typing import Dict, Tuple

import torch

from .transformer_tts import TTSConfig, TransformerTTS

WEIGHT_SEED = 1234
DROPOUT_SEED = 7
# Trained-looking positional scales (P4 makes alpha trainable; 1.0 is only its init).  With alpha = 1
# a random-init decoder's output is dominated by PE and every utterance stops at the same frame;
# 0.25 lets the fed-back frames and the dropout masks matter, so stop times scatter.
ENC_ALPHA = 0.5
DEC_ALPHA = 0.25


def _round_bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def make_model(cfg: TTSConfig | None = None, stop_bias: float = -8.0, seed: int = WEIGHT_SEED) -> TransformerTTS:
    """The canonical oracle model, eval mode."""
    cpu_rng = torch.get_rng_state()
    try:
        torch.manual_seed(seed)
        model = TransformerTTS(cfg)
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            for name, mod in model.named_modules():
                if isinstance(mod, torch.nn.BatchNorm1d):
                    mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
                    mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
                    mod.weight.copy_(1.0 + 0.1 * torch.randn(mod.weight.shape, generator=g))
                    mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
                if isinstance(mod, torch.nn.LayerNorm):
                    mod.weight.copy_(1.0 + 0.1 * torch.randn(mod.weight.shape, generator=g))
                    mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
            model.enc_alpha.fill_(ENC_ALPHA)
            model.dec_alpha.fill_(DEC_ALPHA)
            for name, p in model.named_parameters():
                if p.dim() >= 2:                      # linear / conv / embedding matrices
                    p.copy_(_round_bf16(p))
            model.stop_linear.bias.fill_(stop_bias)
    finally:
        torch.set_rng_state(cpu_rng)
    return model.eval()


def state_dict_digest(sd: Dict[str, torch.Tensor]) -> str:
    """sha256 over the fp32 bytes of every tensor in key order -- pins the canonical weights."""
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().to(torch.float32).contiguous().numpy().tobytes())
    return h.hexdigest()


def make_inputs(B: int, S: int, T: int, data_seed: int, ragged: bool, n_vocab: int = 128, n_mels: int = 80
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """phonemes [B,S] i64 (0 = pad past length), phoneme_lens [B] i32, mels [B,T,80] f32 ~ N(0,1)
    clipped to [-4,4] (zero past length), mel_lens [B] i32.  ragged: lens ~ U[0.6,1.0]*max."""
    g = torch.Generator().manual_seed(data_seed)
    if ragged:
        pl = (torch.rand(B, generator=g) * 0.4 + 0.6) * S
        ml = (torch.rand(B, generator=g) * 0.4 + 0.6) * T
        phoneme_lens = pl.floor().clamp(min=1).to(torch.int32)
        mel_lens = ml.floor().clamp(min=1).to(torch.int32)
        phoneme_lens[0] = S          # keep the padded extent equal to the longest utterance
        mel_lens[0] = T
    else:
        phoneme_lens = torch.full((B,), S, dtype=torch.int32)
        mel_lens = torch.full((B,), T, dtype=torch.int32)
    phonemes = torch.randint(1, n_vocab, (B, S), generator=g, dtype=torch.int64)
    phonemes = phonemes * (torch.arange(S)[None, :] < phoneme_lens[:, None]).to(torch.int64)
    mels = torch.randn(B, T, n_mels, generator=g).clamp(-4.0, 4.0)
    mels = mels * (torch.arange(T)[None, :] < mel_lens[:, None]).to(torch.float32)[..., None]
    return phonemes, phoneme_lens, mels, mel_lens


# BASELINE.json `configs` (index -> shapes); T is frames (teacher-forced length or AR max_len).
CONFIGS = {
    1: dict(B=4, S=100, T=400),
    2: dict(B=1, S=100, T=800),
    3: dict(B=64, S=100, T=800),
    4: dict(B=32, S=100, T=400),
    5: dict(B=16, S=300, T=1600),
}
