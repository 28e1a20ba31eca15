"""Helpers shared by the GPU parity tests."""
import os

import torch


def make_b200_model(oracle_model, cluster_group=0):
    from transformer_tacotron2_b200 import TransformerTTS
    m = TransformerTTS()
    m.load_state_dict(oracle_model.state_dict())
    m.set_option("cluster_group", cluster_group)           # 0 = auto; 1..8 utterances per 8-CTA cluster
    return m


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp(min=1e-12))
