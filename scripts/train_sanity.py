"""Longer training sanity run on one GPU: N optimiser steps on a small fixed synthetic set; prints the loss curve.
Usage (GPU box): python scripts/train_sanity.py [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from transformer_tacotron2_b200.training import Trainer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
model = bench.synthetic_state_dict()
with torch.no_grad():
    model.stop_linear.bias.fill_(0.0)
tr = Trainer(model, lr=3e-4)
g = torch.Generator().manual_seed(1)
B, S, T, NB = 16, 40, 160, 4
data = []
for i in range(NB):
    ph = torch.randint(1, 128, (B, S), generator=g); pl = torch.randint(S // 2, S + 1, (B,), generator=g, dtype=torch.int32); pl[0] = S
    ml = torch.randint(T // 2, T + 1, (B,), generator=g, dtype=torch.int32); ml[0] = T
    # a learnable target: a smooth function of the phoneme sequence (mean phoneme id modulates a sinusoid per mel bin)
    base = ph.float().mean(1, keepdim=True) / 64.0
    t = torch.arange(T)[None, :, None].float(); k = torch.arange(80)[None, None, :].float()
    mel = torch.sin(0.05 * t * (1 + base[:, :, None]) + 0.2 * k) * 1.5
    data.append((ph.cuda(), pl.cuda(), mel.cuda(), ml.cuda()))
losses = []
for s in range(steps):
    tr.lr = 3e-4 * min(1.0, (s + 1) / 20)
    loss = tr.step(*data[s % NB], seed=s)
    if s % 10 == 0 or s == steps - 1:
        losses.append(float(loss)); print(f"step {s:4d}  loss {losses[-1]:.4f}", flush=True)
assert all(l == l for l in losses), "NaN loss"
assert losses[-1] < 0.5 * losses[0], (losses[0], losses[-1])
print("train_sanity ok:", losses[0], "->", losses[-1])
