"""T4/T5: the parity gaps of round 1, closed.

  * free-running AR against the LIVE oracle at the headline length (800 frames, S = 100) and at the long-utterance length
    (1600 frames, S = 300), with the error-growth curve printed;
  * step-locked decoding: GPU step t is fed the ORACLE's frame t-1 (tts_decode_set_frame), so the per-step arithmetic is
    compared without accumulated drift -- tolerances 3-10x tighter than free-running;
  * stop indices / lengths bit-exact on >= 8 searched (data seed, stop bias) cases of B = 8 whose utterances stop at different
    frames INSIDE one cluster group (tests/golden/stop_cases.json, margins asserted);
  * argument validation of the host path before any copy.
pytest -m gpu on a B200."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

from tests.gpu_util import make_b200_model, rel_l2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TOL_AR = 3e-2            # free-running mel rel-L2 (same as test_gpu_parity.py)
TOL_STOP = 5e-2
TOL_LOCKED_MEL = 1e-2    # step-locked: per-frame rel-L2 of mel_before (measured ~3e-3)
TOL_LOCKED_STOP = 2e-2   # step-locked: stop logit max-abs (measured ~5e-3)


@pytest.fixture(scope="module")
def og():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from oracle import synthetic
    o = synthetic.make_model(stop_bias=-8.0)
    return o, make_b200_model(o)


def _growth(ga, ma, name, every):
    T = ma.shape[1]
    pts = []
    for hi in range(every, T + 1, every):
        pts.append(f"{hi}:{rel_l2(ga[:, hi - every:hi], ma[:, hi - every:hi]):.4f}")
    print(f"{name}: rel-L2 per {every}-frame window  " + "  ".join(pts))


@pytest.mark.parametrize("B,S,T,every", [(2, 100, 800, 100), (1, 300, 1600, 200)])
def test_free_running_full_length_vs_live_oracle(og, B, S, T, every):
    from oracle import synthetic
    o, g = og
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ph, pl, _, _ = synthetic.make_inputs(B, S, 8, 90 + T, ragged=True)
    ma, lens, st = o.inference(ph, pl, max_len=T, seed=7)
    ga, gl, gs = (t.cpu() for t in g.inference(ph.cuda(), pl.cuda(), max_len=T, seed=7))
    _growth(ga, ma, f"B={B} S={S} T={T}", every)
    print(f"whole run: mel rel-L2 {rel_l2(ga, ma):.4f}, stop-logit max err {float((gs - st).abs().max()):.4f}")
    assert gl.tolist() == lens.tolist() == [T] * B
    assert rel_l2(ga, ma) < TOL_AR
    assert rel_l2(ga[:, -every:], ma[:, -every:]) < 2 * TOL_AR         # the last window: drift stays bounded
    assert float((gs - st).abs().max()) < TOL_STOP


def test_step_locked_vs_oracle(og):
    """Per-step arithmetic without drift: before GPU step t the oracle's frame t-1 replaces the GPU's own."""
    from oracle import synthetic
    o, g = og
    B, S, T = 4, 40, 96
    ph, pl, _, _ = synthetic.make_inputs(B, S, 8, 123, ragged=True)
    ma, lens, st, mb = o.inference(ph, pl, max_len=T, seed=7, return_before=True)
    lib, hnd = g._ensure_handle(), g._handle
    g.sync_weights()
    ws = g._workspace(B, S, T)
    stream = g._stream()
    phd, pld = ph.cuda(), pl.cuda().int()
    assert lib.tts_decode_begin(hnd, ws.data_ptr(), B, S, T, 7, 0, stream) == 0
    assert lib.tts_encode(hnd, ws.data_ptr(), phd.data_ptr(), pld.data_ptr(), B, S, T, None, stream) == 0
    mbd = mb.cuda()
    gmb = torch.empty(B, T, 80, device="cuda"); gst = torch.empty(B, T, device="cuda")
    for t in range(T):
        if t > 0:                                            # the oracle's previous frame replaces the GPU's own
            fr = mbd[:, t - 1].contiguous()
            assert lib.tts_decode_set_frame(hnd, ws.data_ptr(), t - 1, fr.data_ptr(), stream) == 0
        assert lib.tts_decode_steps(hnd, ws.data_ptr(), 1, stream) == 0
        fo = torch.empty(B, 80, device="cuda"); so = torch.empty(B, device="cuda")
        assert lib.tts_decode_get_frame(hnd, ws.data_ptr(), t, fo.data_ptr(), so.data_ptr(), stream) == 0
        gmb[:, t] = fo; gst[:, t] = so
    torch.cuda.synchronize()
    gmb, gst = gmb.cpu(), gst.cpu()
    per_frame = ((gmb - mb).flatten(2).norm(dim=2) / mb.flatten(2).norm(dim=2).clamp(min=1e-6))      # [B, T]
    print(f"step-locked: mel_before rel-L2 per frame max {float(per_frame.max()):.4f} mean {float(per_frame.mean()):.4f}; "
          f"stop-logit max err {float((gst - st).abs().max()):.4f}")
    assert float(per_frame.max()) < TOL_LOCKED_MEL
    assert float((gst - st).abs().max()) < TOL_LOCKED_STOP


def test_stop_indices_bit_exact_searched_cases():
    """>= 8 cases, B = 8, utterances of one 5-utterance cluster group stop at different frames; lengths bit-exact."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from oracle import synthetic
    z = json.load(open(os.path.join(GOLD, "stop_cases.json")))
    assert len(z["cases"]) >= 8
    models = {}
    worst = 0.0
    for case in z["cases"]:
        bias = case["stop_bias"]
        if bias not in models:
            o = synthetic.make_model(stop_bias=bias)
            models[bias] = (o, make_b200_model(o, cluster_group=5), make_b200_model(o, cluster_group=3))
        o, g5, g3 = models[bias]
        ph, pl, _, _ = synthetic.make_inputs(z["B"], z["S"], 8, case["data_seed"], ragged=True)
        ma, lens, st = o.inference(ph, pl, max_len=z["max_len"], seed=case["seed"])
        assert lens.tolist() == case["lens"], "fixture is stale: regenerate tests/golden/stop_cases.json"
        assert len(set(case["lens"][:5])) >= 2                       # different stop frames inside the first cluster group
        for g in (g5, g3):
            ga, gl, gs = (t.cpu() for t in g.inference(ph.cuda(), pl.cuda(), max_len=z["max_len"], seed=case["seed"]))
            T = min(ga.shape[1], ma.shape[1])
            valid = torch.arange(T)[None, :] < lens[:, None]
            err = float(((gs[:, :T] - st[:, :T]).abs() * valid).max())
            worst = max(worst, err / case["margin"])
            assert case["margin"] > 2 * err, f"case {case}: margin {case['margin']:.4f} <= 2 x logit error {err:.4f}"
            assert gl.tolist() == case["lens"], f"case {case}: gpu lens {gl.tolist()}"
            assert ga.shape[1] == max(case["lens"])
            assert rel_l2(ga[:, :T] * valid[..., None], ma[:, :T] * valid[..., None]) < TOL_AR
    print(f"{len(z['cases'])} stop cases x 2 group sizes bit-exact; worst logit-error / margin ratio {worst:.3f}")


def test_host_path_validates_before_copying(og):
    """tts_infer_host and tts_encode reject S > max_pos / max_len > max_pos with TTS_E_ARG before any copy or launch."""
    o, g = og
    lib, hnd = g._ensure_handle(), g._handle
    g.sync_weights()
    ws = g._workspace(2, 64, 64)
    S_bad = g.cfg.max_pos + 1
    ph = torch.ones(2, S_bad, dtype=torch.int64); pl = torch.full((2,), S_bad, dtype=torch.int32)
    ma = torch.empty(2 * 8 * 80); ml = torch.empty(2, dtype=torch.int32); st = torch.empty(2 * 8); tout = C.c_int(0)
    rc = lib.tts_infer_host(hnd, ws.data_ptr(), ph.data_ptr(), pl.data_ptr(), 2, S_bad, 8, 0, 0, ma.data_ptr(), ml.data_ptr(), st.data_ptr(), C.byref(tout), g._stream())
    assert rc == -1, rc
    rc = lib.tts_infer_host(hnd, ws.data_ptr(), ph.data_ptr(), pl.data_ptr(), 2, 16, g.cfg.max_pos + 1, 0, 0, ma.data_ptr(), ml.data_ptr(), st.data_ptr(), C.byref(tout), g._stream())
    assert rc == -1, rc
    rc = lib.tts_infer_host(hnd, ws.data_ptr(), ph.data_ptr(), pl.data_ptr(), 0, 16, 8, 0, 0, ma.data_ptr(), ml.data_ptr(), st.data_ptr(), C.byref(tout), g._stream())
    assert rc == -1, rc
    phd = ph.cuda()
    rc = lib.tts_encode(hnd, ws.data_ptr(), phd.data_ptr(), pl.cuda().data_ptr(), 2, S_bad, 8, None, g._stream())
    assert rc == -1, rc
    torch.cuda.synchronize()
    # the handle still works afterwards
    ph2 = torch.randint(1, 128, (2, 12)); pl2 = torch.tensor([12, 9], dtype=torch.int32)
    out = g.inference(ph2, pl2, max_len=6, seed=1)
    assert torch.isfinite(out[0]).all()
