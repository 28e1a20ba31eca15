"""SURVEY.md 8(f)-2 / 8(f)-3 on the GPU: ragged batches.  An utterance's result depends on nothing but its own inputs and its
global id (P9), so a batch may be decoded in any order, with per-utterance frame budgets, and with clusters stealing the next
utterance group from a device queue when theirs has stopped -- all bit-identical to the plain call.  pytest -m gpu."""
import pytest
import torch

from tests.gpu_util import make_b200_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from oracle import synthetic
    return make_b200_model(synthetic.make_model(stop_bias=-8.0))


def _inputs(B, S, seed):
    from oracle import synthetic
    ph, pl, _, _ = synthetic.make_inputs(B, S, 8, seed, ragged=True)
    return ph.cuda(), pl.cuda()


def test_sorted_by_length_is_bit_identical(g):
    ph, pl = _inputs(12, 30, 501)
    base = g.inference(ph, pl, max_len=24, seed=7)
    srt = g.inference(ph, pl, max_len=24, seed=7, sort_by_length=True)
    for a, b in zip(base, srt):
        assert torch.equal(a, b)


def test_arbitrary_utterance_ids_key_the_dropout(g):
    ph, pl = _inputs(6, 20, 502)
    ids = [100, 7, 55, 3, 1000, 42]
    out = g.inference(ph, pl, max_len=16, seed=9, utt_ids=ids)
    for i, uid in enumerate(ids):                        # row i == the same utterance decoded alone under its own id
        one = g.inference(ph[i:i + 1], pl[i:i + 1], max_len=16, seed=9, utt_offset=uid)
        assert torch.equal(one[0][0], out[0][i]) and torch.equal(one[2][0], out[2][i])
    plain = g.inference(ph, pl, max_len=16, seed=9)
    assert not torch.equal(plain[0], out[0])             # different ids, different masks


def test_per_utterance_frame_budgets(g):
    ph, pl = _inputs(7, 24, 503)
    budgets = torch.tensor([5, 24, 13, 1, 24, 17, 9], dtype=torch.int32)
    full = g.inference(ph, pl, max_len=24, seed=7)
    out = g.inference(ph, pl, max_len=24, seed=7, max_lens=budgets)
    assert out[1].cpu().tolist() == budgets.tolist()
    assert out[0].shape[1] == 24
    for i, n in enumerate(budgets.tolist()):
        assert torch.equal(out[0][i, :n], full[0][i, :n]) if n == 24 else True           # (postnet context differs at the cut)
        assert torch.equal(out[2][i, :n], full[2][i, :n])                                # stop logits: decoder only
        assert float(out[0][i, n:].abs().max() if n < 24 else 0.0) == 0.0 and float(out[2][i, n:].abs().max() if n < 24 else 0.0) == 0.0
    both = g.inference(ph, pl, max_len=24, seed=7, max_lens=budgets, sort_by_length=True)
    for a, b in zip(out, both):
        assert torch.equal(a, b)


@pytest.mark.parametrize("G", [0, 3])
def test_work_stealing_more_groups_than_clusters(g, G):
    """90 utterances = 18+ groups on <= 15 co-resident clusters with ragged budgets: queue-driven and static assignment agree."""
    from oracle import synthetic
    m = make_b200_model(synthetic.make_model(stop_bias=-8.0), cluster_group=G)
    ph, pl = _inputs(90, 12, 504)
    gen = torch.Generator().manual_seed(5)
    budgets = torch.randint(3, 21, (90,), generator=gen, dtype=torch.int32)
    a = m.inference(ph, pl, max_len=20, seed=3, max_lens=budgets, work_stealing=False)
    b = m.inference(ph, pl, max_len=20, seed=3, max_lens=budgets, work_stealing=True)
    c = m.inference(ph, pl, max_len=20, seed=3, max_lens=budgets, work_stealing=True, sort_by_length=True)
    assert a[1].cpu().tolist() == budgets.tolist()
    for x, y, z in zip(a, b, c):
        assert torch.equal(x, y) and torch.equal(x, z)
