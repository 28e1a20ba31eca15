"""Per-frame error of the decode kernel against the oracle (free-running): finds the first step that goes wrong."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synthetic
from tests.gpu_util import make_b200_model

o = synthetic.make_model(stop_bias=-8.0)
g = make_b200_model(o)
for B, S, T in [(1, 10, 3), (1, 30, 20), (3, 30, 40), (2, 100, 150)]:
    ph, pl, _, _ = synthetic.make_inputs(B, S, 8, 41, ragged=True)
    ma, lens, st = o.inference(ph, pl, max_len=T, seed=7)
    ga, gl, gs = (t.cpu() for t in g.inference(ph.cuda(), pl.cuda(), max_len=T, seed=7))
    err = (gs - st).abs()
    print(f"B={B} S={S} T={T} plens={pl.tolist()}")
    for b in range(B):
        print("  utt", b, "stop-logit err per step:", " ".join(f"{float(e):.3f}" for e in err[b][:24]))
