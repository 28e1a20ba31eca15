"""T0: oracle self-consistency.  CPU only."""
import numpy as np
import torch

from oracle import synthetic, tts_loss


def test_param_count_and_state_dict_keys(oracle_model):
    assert sum(p.numel() for p in oracle_model.parameters()) == 53_011_267
    keys = set(oracle_model.state_dict().keys())
    for k in ("enc_prenet.embed.weight", "enc_prenet.convs.2.bn.running_var", "enc_alpha", "dec_alpha",
              "encoder.layers.5.self_attn.wq.weight", "decoder.layers.0.cross_attn.wo.bias",
              "decoder.layers.5.norm3.weight", "dec_prenet.fc2.weight", "mel_linear.weight",
              "stop_linear.bias", "postnet.convs.4.conv.weight"):
        assert k in keys


def test_inference_equals_teacher_forced_on_own_output(oracle_model_stopping):
    """KV-cache AR decode == re-running forward() on the frames it produced (same Philox masks)."""
    m = oracle_model_stopping
    ph, pl, _, _ = synthetic.make_inputs(3, 24, 8, 11, ragged=True)
    mel_after, lens, stop, mel_before = m.inference(ph, pl, max_len=40, seed=7, return_before=True)
    assert lens.dtype == torch.int32 and mel_after.shape[1] == int(lens.max())
    with torch.no_grad():
        mb2, ma2, st2 = m(ph, pl, mel_before, lens, seed=7)
    assert torch.allclose(mb2, mel_before, atol=2e-4, rtol=1e-4)
    assert torch.allclose(ma2, mel_after, atol=2e-4, rtol=1e-4)
    assert torch.allclose(st2, stop, atol=2e-4, rtol=1e-4)


def test_stop_rule_counts_firing_frame(oracle_model_stopping):
    m = oracle_model_stopping
    ph, pl, _, _ = synthetic.make_inputs(4, 20, 8, 12, ragged=True)
    _, lens, stop, _ = m.inference(ph, pl, max_len=120, seed=7, return_before=True)
    for b in range(4):
        L = int(lens[b])
        fired = (stop[b] > 0).nonzero()
        if len(fired):
            assert L == int(fired[0]) + 1          # firing frame is counted ([TA]:847-851)
            assert (stop[b, :L - 1] <= 0).all()
        else:
            assert L == 120
        assert (stop[b, L:] == 0).all()            # masked past length


def test_batch_composition_independence(oracle_model):
    """P9: an utterance alone == the same utterance inside a padded batch (needs global utt ids)."""
    m = oracle_model
    ph, pl, mels, ml = synthetic.make_inputs(3, 30, 40, 13, ragged=True)
    with torch.no_grad():
        mb, ma, st = m(ph, pl, mels, ml, seed=7)
        b = 2
        S1, T1 = int(pl[b]), int(ml[b])
        mb1, ma1, st1 = m(ph[b:b + 1, :S1], pl[b:b + 1], mels[b:b + 1, :T1], ml[b:b + 1], seed=7, utt_ids=[b])
    assert torch.allclose(mb[b, :T1], mb1[0], atol=2e-5)
    assert torch.allclose(ma[b, :T1], ma1[0], atol=2e-5)
    assert torch.allclose(st[b, :T1], st1[0], atol=2e-5)
    assert (mb[b, T1:] == 0).all() and (ma[b, T1:] == 0).all()


def test_sharded_inference_equals_unsharded(oracle_model_stopping):
    """8(e): splitting the batch (with global utterance ids) reproduces the unsharded result."""
    m = oracle_model_stopping
    ph, pl, _, _ = synthetic.make_inputs(4, 16, 8, 14, ragged=True)
    ma, lens, st = m.inference(ph, pl, max_len=30, seed=7)
    for lo, hi in ((0, 2), (2, 4)):
        ma_s, lens_s, st_s = m.inference(ph[lo:hi], pl[lo:hi], max_len=30, seed=7, utt_ids=list(range(lo, hi)))
        assert torch.equal(lens_s, lens[lo:hi])
        T = ma_s.shape[1]
        assert torch.allclose(ma_s, ma[lo:hi, :T], atol=2e-5)
        assert (ma[lo:hi, T:] == 0).all()


def test_prenet_dropout_is_on_in_eval_and_seeded(oracle_model):
    m = oracle_model
    assert not m.training
    fr = torch.randn(2, 5, 80, generator=torch.Generator().manual_seed(0))
    t, b = np.arange(5), np.arange(2)
    a = m._dec_prenet(fr, 7, t, b)
    assert torch.equal(a, m._dec_prenet(fr, 7, t, b))
    assert not torch.equal(a, m._dec_prenet(fr, 8, t, b))
    h = torch.relu(m.dec_prenet.fc1(fr))
    from oracle import philox as px
    hd = px.dropout_bits(h, 7, px.SITE_DEC_PRENET_FC1, t[None, :], b[:, None])
    kept = hd != 0
    assert torch.allclose(hd[kept], 2.0 * h[kept])          # inverted dropout, scale 2
    assert 0.3 < float((hd == 0).float().mean()) < 0.9


def test_training_mode_runs_and_loss_backward():
    m = synthetic.make_model().train()
    ph, pl, mels, ml = synthetic.make_inputs(2, 12, 16, 15, ragged=True)
    mb, ma, st = m(ph, pl, mels, ml, seed=3)
    loss = tts_loss(mb, ma, st, mels, ml)
    loss.backward()
    assert torch.isfinite(loss)
    assert m.decoder.layers[0].cross_attn.wq.weight.grad.abs().sum() > 0
    assert m.enc_alpha.grad is not None and m.postnet.convs[0].conv.weight.grad.abs().sum() > 0
    # per-utterance masked BatchNorm statistics updated the running stats
    assert int(m.postnet.convs[0].bn.num_batches_tracked) == 1


def test_canonical_weights_are_bf16_representable(oracle_model):
    sd = oracle_model.state_dict()
    for k, v in sd.items():
        if v.dim() >= 2:
            assert torch.equal(v, v.to(torch.bfloat16).to(torch.float32)), k
    assert float(sd["enc_alpha"]) == synthetic.ENC_ALPHA and float(sd["dec_alpha"]) == synthetic.DEC_ALPHA
