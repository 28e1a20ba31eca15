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


def test_train_step_matches_golden(oracle_model):
    """Train-mode forward (all dropout sites, batch-statistics BatchNorm), loss and autograd of the oracle are frozen too."""
    import copy
    from oracle.transformer_tts import tts_loss
    g = np.load(os.path.join(GOLD, "train_small.npz"))
    m = copy.deepcopy(oracle_model).train()
    mels, ml = torch.from_numpy(g["mels"]), torch.from_numpy(g["mel_lens"])
    out = m(torch.from_numpy(g["phonemes"]), torch.from_numpy(g["phoneme_lens"]), mels, ml, seed=int(g["seed"]))
    loss = tts_loss(*out, mels, ml)
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-4
    assert np.allclose(out[1].detach().numpy(), g["mel_after"], atol=2e-4)
    norms = np.array([float(p.grad.norm()) for _, p in m.named_parameters()])
    assert [k for k, _ in m.named_parameters()] == list(g["names"])
    assert np.allclose(norms, g["grad_norms"], rtol=2e-3, atol=1e-7)
    for k, p in m.named_parameters():
        if p.numel() <= 2048:
            assert np.allclose(p.grad.numpy(), g["grad/" + k], atol=2e-5, rtol=2e-3), k
    assert np.allclose(m.postnet.convs[0].bn.running_mean.numpy(), g["bn0_running_mean"], atol=1e-5)
