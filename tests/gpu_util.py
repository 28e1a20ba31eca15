"""Helpers shared by the GPU parity tests."""
import os

import torch


def make_b200_model(oracle_model, persistent=True, cluster=True):
    from transformer_tacotron2_b200 import TransformerTTS
    m = TransformerTTS()
    m.load_state_dict(oracle_model.state_dict())
    if os.environ.get("TTS_FORCE_PER_PHASE") == "1":       # bring-up aid: never launch the persistent kernel
        persistent = False
    m.set_option("decode_persistent", 1 if persistent else 0)
    m.set_option("decode_cluster", 1 if cluster else 0)
    return m


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp(min=1e-12))
