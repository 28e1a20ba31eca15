"""GPU path against the fp32 oracle and against the oracle with bf16-rounded matrix-product operands (oracle/bf16_mode.py):
teacher-forced forward, greedy AR inference, one training step's gradients.  Usage (GPU box): python scripts/diag_bf16_oracle.py"""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synthetic  # noqa: E402
from oracle.bf16_mode import bf16_operands  # noqa: E402
from oracle.transformer_tts import tts_loss  # noqa: E402
from transformer_tacotron2_b200 import TransformerTTS  # noqa: E402
from transformer_tacotron2_b200.training import Trainer  # noqa: E402


def rel(a, b):
    return float((a - b).norm() / b.norm())


import contextlib
for B, S, T in ((4, 30, 60), (8, 100, 200)):
    o = synthetic.make_model(stop_bias=-8.0)
    g = TransformerTTS(device=0)
    g.load_state_dict(o.state_dict())
    ph, pl, mels, ml = synthetic.make_inputs(B, S, T, 77, ragged=True)
    gb, ga, gs = (t.cpu() for t in g(ph, pl, mels, ml, seed=7))
    gia, gil, gis = g.inference(ph, pl, max_len=min(T, 80), seed=7)
    tr = Trainer(g)
    loss = float(tr.forward_backward(ph, pl, mels, ml, seed=7))
    grads = tr.grads()
    for name, ctx in (("fp32 oracle", contextlib.nullcontext), ("bf16-operand oracle", bf16_operands)):
        with ctx():
            with torch.no_grad():
                mb, ma, st = o(ph, pl, mels, ml, seed=7)
                ia, il, is_ = o.inference(ph, pl, max_len=min(T, 80), seed=7)
            ot = copy.deepcopy(o).train()
            lref = tts_loss(*ot(ph, pl, mels, ml, seed=7), mels, ml)
            lref.backward()
        per = {k: rel(grads[k], p.grad) for k, p in ot.named_parameters() if float(p.grad.norm()) > 0}
        num = sum(float((grads[k] - p.grad).norm() ** 2) for k, p in ot.named_parameters()) ** 0.5
        den = sum(float(p.grad.norm() ** 2) for p in ot.parameters()) ** 0.5
        worst = sorted(per.items(), key=lambda kv: -kv[1])[:3]
        print(f"B{B} S{S} T{T} vs {name:20s}: forward mel_before {rel(gb, mb):.5f} mel_after {rel(ga, ma):.5f} stop max-abs {float((gs - st).abs().max()):.5f} | "
              f"AR mel {rel(gia, ia):.5f} lens equal {gil.tolist() == il.tolist()} | loss rel {abs(loss - float(lref)) / abs(float(lref)):.2e} "
              f"grad global {num / den:.4f} worst {[(k, round(v, 4)) for k, v in worst]}", flush=True)
    del tr, g
