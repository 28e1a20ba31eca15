import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_state_dict, synthetic_inputs
from transformer_tacotron2_b200 import TransformerTTS
m = TransformerTTS(); m.load_state_dict(synthetic_state_dict().state_dict())
ph, pl = synthetic_inputs(2, 8, 5)
m.set_option('debug_mode', int(os.environ.get('TTS_DBG', '0')))
a, l, s = m.inference(ph.cuda(), pl.cuda(), max_len=3, seed=7)
torch.cuda.synchronize()
print("ok", a.shape, l.tolist(), bool(torch.isfinite(a).all()))
