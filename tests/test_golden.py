"""Oracle vs the committed golden fixtures (tests/golden/make_golden.py made them)."""
import os

import numpy as np
import torch

from oracle import synthetic

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_canonical_weight_digest(oracle_model):
    g = np.load(os.path.join(GOLD, "forward_small.npz"))
    assert synthetic.state_dict_digest(oracle_model.state_dict()) == str(g["digest"])


def test_forward_matches_golden(oracle_model):
    g = np.load(os.path.join(GOLD, "forward_small.npz"))
    with torch.no_grad():
        mb, ma, st = oracle_model(torch.from_numpy(g["phonemes"]), torch.from_numpy(g["phoneme_lens"]),
                                  torch.from_numpy(g["mels"]), torch.from_numpy(g["mel_lens"]), seed=7)
    assert np.allclose(mb.numpy(), g["mel_before"], atol=1e-4)
    assert np.allclose(ma.numpy(), g["mel_after"], atol=1e-4)
    assert np.allclose(st.numpy(), g["stop_logits"], atol=1e-4)


def test_inference_matches_golden():
    g = np.load(os.path.join(GOLD, "inference_small.npz"))
    model = synthetic.make_model(stop_bias=float(g["stop_bias"]))
    ma, lens, st = model.inference(torch.from_numpy(g["phonemes"]),
                                                   torch.from_numpy(g["phoneme_lens"]), max_len=48, seed=7)
    assert lens.tolist() == g["mel_lens"].tolist()
    assert np.allclose(ma.numpy(), g["mel_after"], atol=2e-4)
    assert np.allclose(st.numpy(), g["stop_logits"], atol=2e-4)
    assert len(set(lens.tolist())) == 3 and int(lens.max()) < 48          # ragged stops, all before max_len
    T = st.shape[1]
    valid = torch.arange(T)[None, :] < lens[:, None]
    assert float(st.abs()[valid].min()) > 0.04                             # the searched stop margin
