"""checkpoint naming interchange and the vocoder layout (CPU)."""
import pytest
import torch

from oracle import synthetic
from transformer_tacotron2_b200 import interchange as ic


@pytest.fixture(scope="module")
def ref_sd():
    return synthetic.make_model(stop_bias=-8.0).state_dict()


def test_espnet_round_trip(ref_sd):
    esp = ic.to_espnet_state_dict(ref_sd)
    assert "encoder.encoders.0.self_attn.linear_q.weight" in esp and "decoder.decoders.5.src_attn.linear_out.bias" in esp
    assert "feat_out.weight" in esp and "prob_out.bias" in esp and "postnet.postnet.4.1.running_var" in esp
    assert esp["encoder.embed.1.alpha"].shape == (1,)
    esp = {"tts." + k: v for k, v in esp.items()}
    esp["tts.encoder.embed.1.pe"] = torch.zeros(1, 10, 512)            # positional tables are recomputed, not imported
    back = ic.from_espnet_state_dict(esp, ref_sd)
    assert set(back) == set(ref_sd)
    for k in ref_sd:
        assert torch.equal(back[k].to(ref_sd[k].dtype), ref_sd[k]), k
    m = synthetic.make_model(stop_bias=0.0)
    m.load_state_dict(back)                                            # loads into the oracle (same keys as the B200 module)


def test_espnet_convs_without_bias(ref_sd):
    esp = {k: v for k, v in ic.to_espnet_state_dict(ref_sd).items() if not (k.endswith(".0.bias") and (".convs." in k or "postnet.postnet" in k))}
    back = ic.from_espnet_state_dict(esp, ref_sd)
    assert float(back["postnet.convs.0.conv.bias"].abs().max()) == 0.0


def test_import_refuses_incompatible_checkpoints(ref_sd):
    esp = ic.to_espnet_state_dict(ref_sd)
    with pytest.raises(ic.InterchangeError, match="normalize_before"):
        ic.from_espnet_state_dict({**esp, "encoder.after_norm.weight": torch.ones(512)}, ref_sd)
    with pytest.raises(ic.InterchangeError, match="shape"):
        ic.from_espnet_state_dict({**esp, "feat_out.weight": torch.zeros(160, 512)}, ref_sd)     # reduction factor 2
    with pytest.raises(ic.InterchangeError, match="lacks"):
        ic.from_espnet_state_dict({k: v for k, v in esp.items() if not k.startswith("decoder.decoders.5.")}, ref_sd)
    with pytest.raises(ic.InterchangeError, match="unrecognised"):
        ic.from_espnet_state_dict({**esp, "spk_embed.weight": torch.zeros(3)}, ref_sd)
    with pytest.raises(ic.InterchangeError, match="more layers"):
        ic.from_espnet_state_dict({**esp, "encoder.encoders.6.norm1.weight": torch.ones(512)}, ref_sd)


def test_mel_for_vocoder_layout():
    mel = torch.arange(2 * 5 * 80, dtype=torch.float32).view(2, 5, 80)
    lens = torch.tensor([5, 3], dtype=torch.int32)
    v = ic.mel_for_vocoder(mel, lens, mean=torch.full((80,), 1.0), std=torch.full((80,), 2.0), pad_value=-11.5)
    assert v.shape == (2, 80, 5) and v.is_contiguous() and v.dtype == torch.float32
    assert torch.equal(v[0, :, 4], mel[0, 4] * 2 + 1)
    assert float(v[1, :, 3:].max()) == -11.5 and float(v[1, :, 3:].min()) == -11.5
