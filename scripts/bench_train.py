"""Train-step throughput (utt/s) of the B200 path at the BASELINE config-3 shapes (B = 32 per GPU, S = 100, T = 400 / 800).
Usage (GPU box): python scripts/bench_train.py [B S T]  -> gpurun_out/train_bench.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from transformer_tacotron2_b200.training import Trainer  # noqa: E402


def run(B, S, T, steps=int(os.environ.get("STEPS", "5"))):
    model = bench.synthetic_state_dict()
    tr = Trainer(model, lr=1e-4)
    g = torch.Generator().manual_seed(B + T)
    ph = torch.randint(1, 128, (B, S), generator=g).cuda(); pl = torch.full((B,), S, dtype=torch.int32).cuda()
    mel = torch.randn(B, T, 80, generator=g).clamp(-4, 4).cuda(); ml = torch.full((B,), T, dtype=torch.int32).cuda()
    for i in range(2):
        tr.step(ph, pl, mel, ml, seed=i)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for i in range(steps):
        tr.forward_backward(ph, pl, mel, ml, seed=10 + i)
    e[1].record()
    for i in range(steps):
        tr.step(ph, pl, mel, ml, seed=20 + i)
    e[2].record(); torch.cuda.synchronize()
    fb, full = e[0].elapsed_time(e[1]) / steps, e[1].elapsed_time(e[2]) / steps
    print(f"B={B} S={S} T={T}: forward+backward {fb:.2f} ms, full step (with Adam + repack) {full:.2f} ms -> {B / full * 1e3:.1f} utt/s, loss {float(tr._loss):.4f}")
    return dict(B=B, S=S, T=T, fwd_bwd_ms=fb, step_ms=full, utt_per_s=B / full * 1e3)


if __name__ == "__main__":
    if len(sys.argv) > 3:
        shapes = [tuple(int(x) for x in sys.argv[1:4])]
    else:
        shapes = [(32, 100, 400), (32, 100, 800)]
    out = [run(*s) for s in shapes]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "train_bench.json"), "w"), indent=1)
