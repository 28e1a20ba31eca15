"""Decode time per step (CUDA events around the persistent decode kernel).  Usage (GPU box):
python scripts/time_decode.py [B] [T] [S] [key=value,...]   (options go to tts_set_option)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_state_dict, synthetic_inputs  # noqa: E402
from transformer_tacotron2_b200 import TransformerTTS  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 800
S = int(sys.argv[3]) if len(sys.argv) > 3 else 100
m = TransformerTTS()
m.load_state_dict(synthetic_state_dict().state_dict())
for kv in (sys.argv[4].split(",") if len(sys.argv) > 4 else []):
    k, v = kv.split("=")
    m.set_option(k, int(v))
ph, pl = synthetic_inputs(B, S, 103)
ph, pl = ph.cuda(), pl.cuda()
m.profile_events = True
for _ in range(4):
    m.inference(ph, pl, max_len=T, seed=7)
print(f"B={B} T={T} S={S}: decode {min(m.decode_ms):8.2f} ms -> {1e3 * min(m.decode_ms) / T:6.1f} us/step", flush=True)
