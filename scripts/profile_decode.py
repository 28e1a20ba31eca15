"""Per-phase timing of the persistent decode kernel from in-kernel %globaltimer stamps.
Usage (GPU box): python scripts/profile_decode.py [B] [T] [S]   -> prints a table, writes gpurun_out/phase_times.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_state_dict, synthetic_inputs, decode_bytes  # noqa: E402
from transformer_tacotron2_b200 import TransformerTTS  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 800
S = int(sys.argv[3]) if len(sys.argv) > 3 else 100
names = ["fc1", "fc2", "proj"] + [f"L{l}.{n}" for l in range(6) for n in ("qkv", "self_attn", "o", "cq", "cross_attn", "o2", "ffn1", "ffn2")] + ["head"]
m = TransformerTTS()
m.load_state_dict(synthetic_state_dict().state_dict())
m.set_option("decode_timestamps", 1)
ph, pl = synthetic_inputs(B, S, 103)
ph, pl = ph.cuda(), pl.cuda()
for _ in range(2):
    m.inference(ph, pl, max_len=T, seed=7)
ts = m.phase_timestamps(T).double()                    # [T, 52] ns, stamp taken after each phase's barrier
flat = ts.view(-1)
d = (flat[1:] - flat[:-1]).view(-1)
d = torch.cat([d[:1] * 0, d]).view(T, len(names)) / 1e3   # us per phase (first phase of step 0 unknown -> 0)
step_us = d.sum(1)
groups = {}
for i, n in enumerate(names):
    key = n.split(".")[-1]
    groups.setdefault(key, []).append(i)
out = {"B": B, "T": T, "S": S, "us_per_step_mean": float(step_us[1:].mean()), "windows": {}}
for lo, hi in ((1, 50), (T // 2 - 25, T // 2 + 25), (T - 50, T)):
    w = d[lo:hi]
    row = {k: float(w[:, idx].mean()) for k, idx in groups.items()}      # mean us per phase instance
    tot = {k: float(w[:, idx].sum(1).mean()) for k, idx in groups.items()}  # us per step for the whole group
    out["windows"][f"{lo}-{hi}"] = {"per_phase_us": row, "per_step_us": tot, "step_us": float(w.sum(1).mean())}
    print(f"steps {lo}-{hi}: {w.sum(1).mean():.1f} us/step")
    for k in row:
        print(f"   {k:11s} {row[k]:7.2f} us/phase x{len(groups[k]):2d} = {tot[k]:7.1f} us/step")
a = m.all_stamps.double()
if (a[:, 52] > 0).any():
    for name, w in (("t~400", a[T // 2 - 100: T // 2 + 100]), ("t~750", a[T - 100: T - 5])):
        d = lambda i, j: float((w[:, i] - w[:, j]).mean() / 1e3)
        print(f"[{name}] layer 0, us:")
        print("  self-attn: loop %.2f | partial bar %.2f | merge+push %.2f | gather %.2f" % (d(61, 3), d(62, 61), d(63, 62), d(4, 63)))
        if (w[:, 83] > 0).any():
            print("  warp-0 attention loop (durations): self wait %.2f | compute %.2f | release+refill %.2f    cross wait %.2f | compute %.2f | release+refill %.2f" % tuple(
                float(w[:, i].mean() / 1965.0) for i in (82, 83, 84, 85, 86, 87)))   # SM clock cycles at 1965 MHz
            print("  warp-12 (pair 4) loop (durations)  : self wait %.2f | compute %.2f | release+refill %.2f    cross wait %.2f | compute %.2f | release+refill %.2f" % tuple(
                float(w[:, i].mean() / 1965.0) for i in (88, 89, 90, 91, 92, 93)))
        print("  O-proj   : gemm+epi %.2f | push %.2f | gather %.2f | LN %.2f" % (d(52, 4), d(53, 52), d(54, 53), d(5, 54)))
        print("  cross    : q2 gemm+bar %.2f | loop %.2f | partial bar %.2f | merge+push %.2f | gather %.2f" % (d(6, 5), d(64, 6), d(65, 64), d(66, 65), d(7, 66)))
        print("  warp-0 GEMM (wait for stage | MMA loop): O-proj %.2f | %.2f   FFN1 %.2f | %.2f   FFN2 %.2f | %.2f" % (d(74, 73), d(75, 74), d(77, 76), d(78, 77), d(80, 79), d(81, 80)))
        if (w[:, 94] > 0).any():
            print("  y3 detail: rs gather done -> reduced %.2f | -> pushed %.2f" % (d(94, 58), d(59, 94)))
        print("  FFN      : ffn1 %.2f | ffn2 gemm (warp 0) %.2f | rs push (warp 0) %.2f | rs gather %.2f | reduce+y3 push %.2f | y3 gather %.2f | LN %.2f" % (
            d(9, 8), d(55, 9), d(57, 55), d(58, 57), d(59, 58), d(60, 59), d(10, 60)))
print("mean us/step", out["us_per_step_mean"], " roofline us/step", decode_bytes(B, T, S) / T / 6468.6e3)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/phase_times.json", "w"), indent=1)
m.set_option("print_info", 1)
if os.environ.get("TTS_GROUPS"):
    import time
    for G in [int(x) for x in os.environ["TTS_GROUPS"].split(",")]:
        m.set_option("cluster_group", G); m.set_option("decode_timestamps", 0)
        m.profile_events = True; m.decode_ms.clear()
        for _ in range(3):
            m.inference(ph, pl, max_len=T, seed=7)
        print(f"cluster_group={G}: decode {min(m.decode_ms):.2f} ms -> {1e3*min(m.decode_ms)/T:.1f} us/step")

