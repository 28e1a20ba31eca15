"""Regenerates tests/golden/*.npz from the oracle (run from the repo root: python tests/golden/make_golden.py).

The reference repository has no golden vectors (it has no code: /root/reference/README.md:1-3), so
these fixtures are outputs of the oracle on the canonical synthetic weights; they freeze the
oracle's behaviour (tests/test_golden.py, CPU) and are what the CUDA path is compared against on
the GPU box (tests/test_gpu_parity.py), where the oracle is also re-run live.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import synthetic  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

FORWARD_CASE = dict(B=2, S=12, T=20, data_seed=201, seed=7, stop_bias=-8.0)
# (data_seed, stop_bias) searched so that the oracle's minimum |stop logit| over the valid frames (the
# "stop margin", 0.046 here) is well above the bf16 path's logit error (~0.006): only then is
# "stop indices bit-exact" a meaningful requirement.
INFER_CASE = dict(B=3, S=16, max_len=48, data_seed=212, seed=7, stop_bias=-0.505)
# one training step (train-mode forward with every dropout site on, loss P13, autograd): loss, train-mode outputs, the small
# parameter gradients in full and the L2 norm of every gradient tensor
TRAIN_CASE = dict(B=3, S=14, T=24, data_seed=233, seed=11, stop_bias=-8.0)


def main():
    torch.set_num_threads(1)
    c = FORWARD_CASE
    m = synthetic.make_model(stop_bias=c["stop_bias"])
    digest = synthetic.state_dict_digest(m.state_dict())
    ph, pl, mels, ml = synthetic.make_inputs(c["B"], c["S"], c["T"], c["data_seed"], ragged=True)
    with torch.no_grad():
        mb, ma, st = m(ph, pl, mels, ml, seed=c["seed"])
        mem = m.encode(ph, pl)
    np.savez_compressed(os.path.join(HERE, "forward_small.npz"), digest=np.array(digest),
                        phonemes=ph.numpy(), phoneme_lens=pl.numpy(), mels=mels.numpy(), mel_lens=ml.numpy(),
                        memory=mem.numpy(), mel_before=mb.numpy(), mel_after=ma.numpy(), stop_logits=st.numpy())
    c = INFER_CASE
    m = synthetic.make_model(stop_bias=c["stop_bias"])
    ph, pl, _, _ = synthetic.make_inputs(c["B"], c["S"], 8, c["data_seed"], ragged=True)
    ma, lens, st, mb = m.inference(ph, pl, max_len=c["max_len"], seed=c["seed"], return_before=True)
    np.savez_compressed(os.path.join(HERE, "inference_small.npz"),
                        phonemes=ph.numpy(), phoneme_lens=pl.numpy(), mel_after=ma.numpy(), mel_before=mb.numpy(),
                        mel_lens=lens.numpy(), stop_logits=st.numpy(), stop_bias=np.array(c["stop_bias"]),
                        max_len=np.array(c["max_len"]), seed=np.array(c["seed"]))
    import copy
    from oracle.transformer_tts import tts_loss
    c = TRAIN_CASE
    m = copy.deepcopy(synthetic.make_model(stop_bias=c["stop_bias"])).train()
    ph, pl, mels, ml = synthetic.make_inputs(c["B"], c["S"], c["T"], c["data_seed"], ragged=True)
    out = m(ph, pl, mels, ml, seed=c["seed"])
    loss = tts_loss(*out, mels, ml)
    loss.backward()
    names = [k for k, _ in m.named_parameters()]
    norms = np.array([float(p.grad.norm()) for _, p in m.named_parameters()], dtype=np.float64)
    small = {("grad/" + k): p.grad.numpy() for k, p in m.named_parameters() if p.numel() <= 2048}
    np.savez_compressed(os.path.join(HERE, "train_small.npz"), phonemes=ph.numpy(), phoneme_lens=pl.numpy(), mels=mels.numpy(),
                        mel_lens=ml.numpy(), seed=np.array(c["seed"]), loss=np.array(float(loss)), names=np.array(names), grad_norms=norms,
                        mel_before=out[0].detach().numpy(), mel_after=out[1].detach().numpy(), stop_logits=out[2].detach().numpy(),
                        bn0_running_mean=m.postnet.convs[0].bn.running_mean.numpy(), **small)
    print("digest", digest, "lens", lens.tolist(), "train loss", float(loss))


if __name__ == "__main__":
    main()
