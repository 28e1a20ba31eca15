"""Per-tensor gradient error of the B200 training step against the oracle's autograd (diagnostic; GPU box)."""
import copy, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synthetic
from oracle.transformer_tts import tts_loss
from transformer_tacotron2_b200 import TransformerTTS
from transformer_tacotron2_b200.training import Trainer

B, S, T, ragged = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (3, 12, 20, 1))]
p_res = float(os.environ.get("P_RES", "0.1"))
om = synthetic.make_model(stop_bias=-8.0)
object.__setattr__(om.cfg, "p_residual", p_res)
inputs = synthetic.make_inputs(B, S, T, 900 + B, bool(ragged))
m = copy.deepcopy(om).train()
out = m(*inputs, seed=7); loss = tts_loss(*out, inputs[2], inputs[3]); loss.backward()
model = TransformerTTS(); model.load_state_dict(om.state_dict())
tr = Trainer(model, p_residual=p_res)
l = tr.forward_backward(*inputs, seed=7)
print("loss", float(l), float(loss))
g = tr.grads()
tot = torch.cat([p.grad.flatten() for p in m.parameters()]).norm()
for k, p in m.named_parameters():
    e = float((g[k] - p.grad).norm() / p.grad.norm().clamp(min=1e-20))
    print(f"{k:48s} |g|/tot {float(p.grad.norm() / tot):9.2e}  rel err {e:9.3e}")
num = sum(float((g[k] - p.grad).norm() ** 2) for k, p in m.named_parameters()) ** 0.5
print("global rel err", num / float(tot))
