"""Bring-up check: cluster decode kernel vs the grid-barrier kernel on the same inputs (GPU box)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_state_dict, synthetic_inputs  # noqa: E402
from transformer_tacotron2_b200 import TransformerTTS  # noqa: E402

sd = synthetic_state_dict().state_dict()
mc = TransformerTTS(); mc.load_state_dict(sd)
mg = TransformerTTS(); mg.load_state_dict(sd); mg.set_option("decode_cluster", 0)
for (B, S, T) in [(3, 16, 12), (8, 20, 40), (11, 24, 30), (1, 10, 20), (64, 100, 150)]:
    ph, pl = synthetic_inputs(B, S, 5)
    pl = torch.randint(max(1, S // 2), S + 1, (B,), dtype=torch.int32); pl[0] = S
    t0 = time.time()
    a1, l1, s1 = (x.cpu() for x in mc.inference(ph.cuda(), pl.cuda(), max_len=T, seed=7))
    t1 = time.time()
    a2, l2, s2 = (x.cpu() for x in mg.inference(ph.cuda(), pl.cuda(), max_len=T, seed=7))
    rel = float((a1 - a2).norm() / a2.norm())
    print(f"B={B} S={S} T={T}: cluster {t1 - t0:.3f}s lens {l1.tolist()[:4]} vs {l2.tolist()[:4]} mel rel-L2 {rel:.5f} "
          f"stop max-abs {float((s1 - s2).abs().max()):.5f} finite {bool(torch.isfinite(a1).all())}", flush=True)
