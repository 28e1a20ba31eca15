"""ORACLE (test infrastructure, not product code) -- PyTorch-eager CPU fp32 Transformer-TTS.

PARITY UNPINNED BY THE REFERENCE: keonlee9420/Transformer-tacotron2 ships no code, tests or golden
vectors (/root/reference/README.md:1-3 is its whole content; README.md:3 links Li et al., "Neural
Speech Synthesis with Transformer Network", AAAI 2019).  BASELINE.json `north_star` therefore
makes THIS module both the API surface to keep and the correctness baseline.  Its sub-modules are
pinned instead against the two independent upstream implementations installed in the image
(tests/test_oracle_upstream.py):

  [TT] torch/nn/modules/transformer.py  -- TransformerEncoderLayer / TransformerDecoderLayer,
       post-LN order (lines 952-956, 1144-1153), eps 1e-5, d_ff 2048 (lines 102-113)
  [TA] torchaudio/models/tacotron2.py   -- _Prenet (258-285, dropout training=True at 284),
       _Postnet (288-346), _Encoder.convolutions (371-385, 407-408), stop bookkeeping (847-853)

Every under-specified choice is frozen here as SURVEY.md section 8-P lists it (P1..P18); the tags
below point at those rows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import philox as px


@dataclass(frozen=True)
class TTSConfig:
    n_vocab: int = 128
    d_model: int = 512
    n_heads: int = 8
    n_enc_layers: int = 6
    n_dec_layers: int = 6
    d_ff: int = 2048            # P2
    n_mels: int = 80
    d_prenet: int = 256
    enc_conv_layers: int = 3
    conv_kernel: int = 5
    postnet_channels: int = 512
    postnet_layers: int = 5
    max_pos: int = 2048         # P4: PE table length
    ln_eps: float = 1e-5
    bn_eps: float = 1e-5
    p_prenet: float = 0.5       # P7 (always on), P6/P11 (train only)
    p_residual: float = 0.1     # P12 (attention-probability dropout is 0)
    stop_pos_weight: float = 5.0  # P13

    def to_dict(self):
        return asdict(self)


def sinusoid_table(n_pos: int, d: int) -> torch.Tensor:
    """P4: PE[pos, 2i] = sin(pos / 10000^(2i/d)), PE[pos, 2i+1] = cos(same); fp32, computed in
    float64 then rounded so that every backend can reproduce it bit-exactly from this definition."""
    pos = torch.arange(n_pos, dtype=torch.float64)[:, None]
    i2 = torch.arange(0, d, 2, dtype=torch.float64)[None, :]
    ang = pos / torch.pow(torch.tensor(10000.0, dtype=torch.float64), i2 / d)
    pe = torch.zeros(n_pos, d, dtype=torch.float64)
    pe[:, 0::2] = torch.sin(ang)
    pe[:, 1::2] = torch.cos(ang)
    return pe.to(torch.float32)


def length_mask(lens: torch.Tensor, n: int) -> torch.Tensor:
    """[B, n] bool, True at valid positions."""
    return torch.arange(n)[None, :] < lens.to(torch.int64)[:, None]


class MultiHeadAttention(nn.Module):
    """P3: 8 x 64, scale 1/8, biases on q/k/v/o, separate projection matrices."""

    def __init__(self, d: int, h: int):
        super().__init__()
        self.h, self.dh = h, d // h
        self.wq, self.wk, self.wv, self.wo = (nn.Linear(d, d) for _ in range(4))

    def split(self, x):                     # [B, L, d] -> [B, h, L, dh]
        B, L, _ = x.shape
        return x.view(B, L, self.h, self.dh).transpose(1, 2)

    def attend(self, q, k, v, mask):        # mask: bool broadcastable to [B,h,Lq,Lk], True = keep
        s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(self.dh)
        if mask is not None:
            s = s.masked_fill(~mask, float("-inf"))
        p = torch.softmax(s, dim=-1)
        o = torch.matmul(p, v)              # [B,h,Lq,dh]
        B, _, Lq, _ = o.shape
        return self.wo(o.transpose(1, 2).reshape(B, Lq, self.h * self.dh))

    def forward(self, xq, xkv, mask):
        return self.attend(self.split(self.wq(xq)), self.split(self.wk(xkv)), self.split(self.wv(xkv)), mask)


class FeedForward(nn.Module):
    def __init__(self, d: int, d_ff: int):
        super().__init__()
        self.w1, self.w2 = nn.Linear(d, d_ff), nn.Linear(d_ff, d)

    def forward(self, x):
        return self.w2(F.relu(self.w1(x)))


class EncoderLayer(nn.Module):
    """P1: post-LN, [TT]:952-956."""

    def __init__(self, c: TTSConfig):
        super().__init__()
        self.self_attn = MultiHeadAttention(c.d_model, c.n_heads)
        self.norm1 = nn.LayerNorm(c.d_model, eps=c.ln_eps)
        self.ffn = FeedForward(c.d_model, c.d_ff)
        self.norm2 = nn.LayerNorm(c.d_model, eps=c.ln_eps)


class DecoderLayer(nn.Module):
    """P1: post-LN, [TT]:1144-1153."""

    def __init__(self, c: TTSConfig):
        super().__init__()
        self.self_attn = MultiHeadAttention(c.d_model, c.n_heads)
        self.norm1 = nn.LayerNorm(c.d_model, eps=c.ln_eps)
        self.cross_attn = MultiHeadAttention(c.d_model, c.n_heads)
        self.norm2 = nn.LayerNorm(c.d_model, eps=c.ln_eps)
        self.ffn = FeedForward(c.d_model, c.d_ff)
        self.norm3 = nn.LayerNorm(c.d_model, eps=c.ln_eps)


class ConvBN(nn.Module):
    def __init__(self, cin: int, cout: int, k: int, eps: float):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, k, padding=(k - 1) // 2)
        self.bn = nn.BatchNorm1d(cout, eps=eps, momentum=0.1)


def masked_batchnorm(bn: nn.BatchNorm1d, x: torch.Tensor, mask: torch.Tensor, training: bool) -> torch.Tensor:
    """P5. x [B,C,L], mask [B,1,L] float.  Inference: running stats.  Training: batch statistics
    over the VALID positions only (biased variance for normalisation, unbiased for the running
    update, momentum 0.1), no SyncBN."""
    if not training:
        return F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps)
    n = mask.sum().clamp(min=1.0)
    mean = (x * mask).sum(dim=(0, 2)) / n
    var = (((x - mean[None, :, None]) ** 2) * mask).sum(dim=(0, 2)) / n
    with torch.no_grad():
        bn.running_mean.mul_(1 - bn.momentum).add_(bn.momentum * mean)
        bn.running_var.mul_(1 - bn.momentum).add_(bn.momentum * var * n / (n - 1).clamp(min=1.0))
        bn.num_batches_tracked += 1
    xh = (x - mean[None, :, None]) * torch.rsqrt(var[None, :, None] + bn.eps)
    return xh * bn.weight[None, :, None] + bn.bias[None, :, None]


class EncoderPrenet(nn.Module):
    """P6: Embedding(V,512,pad 0) -> 3 x [Conv1d k5 -> BN -> ReLU -> Dropout(.5 train)] -> Linear."""

    def __init__(self, c: TTSConfig):
        super().__init__()
        self.embed = nn.Embedding(c.n_vocab, c.d_model, padding_idx=0)
        self.convs = nn.ModuleList(ConvBN(c.d_model, c.d_model, c.conv_kernel, c.bn_eps) for _ in range(c.enc_conv_layers))
        self.proj = nn.Linear(c.d_model, c.d_model)


class DecoderPrenet(nn.Module):
    """P7: Linear(80,256) ReLU Drop -> Linear(256,256) ReLU Drop -> Linear(256,512); biases on."""

    def __init__(self, c: TTSConfig):
        super().__init__()
        self.fc1 = nn.Linear(c.n_mels, c.d_prenet)
        self.fc2 = nn.Linear(c.d_prenet, c.d_prenet)
        self.proj = nn.Linear(c.d_prenet, c.d_model)


class Postnet(nn.Module):
    """P11: 80 -> 512 -> 512 -> 512 -> 512 -> 80, BN each, tanh on the first four ([TA]:306-346)."""

    def __init__(self, c: TTSConfig):
        super().__init__()
        ch = [c.n_mels] + [c.postnet_channels] * (c.postnet_layers - 1) + [c.n_mels]
        self.convs = nn.ModuleList(ConvBN(ch[i], ch[i + 1], c.conv_kernel, c.bn_eps) for i in range(c.postnet_layers))


class _Stack(nn.Module):
    def __init__(self, layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)


class TransformerTTS(nn.Module):
    """The API surface: `forward` teacher-forced, `inference` greedy AR."""

    def __init__(self, cfg: Optional[TTSConfig] = None):
        super().__init__()
        self.cfg = c = cfg or TTSConfig()
        self.enc_prenet = EncoderPrenet(c)
        self.enc_alpha = nn.Parameter(torch.ones(()))     # P4
        self.dec_alpha = nn.Parameter(torch.ones(()))
        self.encoder = _Stack(EncoderLayer(c) for _ in range(c.n_enc_layers))
        self.dec_prenet = DecoderPrenet(c)
        self.decoder = _Stack(DecoderLayer(c) for _ in range(c.n_dec_layers))
        self.mel_linear = nn.Linear(c.d_model, c.n_mels)
        self.stop_linear = nn.Linear(c.d_model, 1)
        self.postnet = Postnet(c)
        self.register_buffer("pe", sinusoid_table(c.max_pos, c.d_model), persistent=False)

    # ------------------------------------------------------------------ dropout helpers
    def _drop_res(self, x, seed, site, t, b):
        """P12 residual dropout (train only). x [B,L,d]; t [L]; b [B]."""
        if not self.training or self.cfg.p_residual == 0.0:
            return x
        return px.dropout_words(x, seed, site, t[None, :], b[:, None], self.cfg.p_residual)

    # ------------------------------------------------------------------ encoder
    def encode(self, phonemes, phoneme_lens, seed: int = 0, utt_ids=None) -> torch.Tensor:
        c = self.cfg
        B, S = phonemes.shape
        b_ids = np.arange(B) if utt_ids is None else np.asarray(utt_ids)
        t_ids = np.arange(S)
        valid = length_mask(phoneme_lens, S)                        # [B,S]
        m = valid[:, None, :].to(torch.float32)                     # [B,1,S]
        x = self.enc_prenet.embed(phonemes).transpose(1, 2) * m     # [B,512,S]; P9 zero past length
        for i, cb in enumerate(self.enc_prenet.convs):
            x = F.relu(masked_batchnorm(cb.bn, cb.conv(x), m, self.training))
            if self.training:
                x = px.dropout_bits(x.transpose(1, 2), seed, px.SITE_ENC_PRENET_CONV0 + i,
                                    t_ids[None, :], b_ids[:, None]).transpose(1, 2)
            x = x * m                                               # P9
        x = self.enc_prenet.proj(x.transpose(1, 2))                 # [B,S,512]
        x = x + self.enc_alpha * self.pe[:S][None]                  # P4
        x = self._drop_res(x, seed, px.SITE_ENC_PE, t_ids, b_ids)
        kmask = valid[:, None, None, :]                             # key padding
        for l, layer in enumerate(self.encoder.layers):
            a = layer.self_attn(x, x, kmask)
            x = layer.norm1(x + self._drop_res(a, seed, px.SITE_ENC_LAYER0 + 2 * l, t_ids, b_ids))
            f = layer.ffn(x)
            x = layer.norm2(x + self._drop_res(f, seed, px.SITE_ENC_LAYER0 + 2 * l + 1, t_ids, b_ids))
        return x                                                    # memory [B,S,512]

    # ------------------------------------------------------------------ decoder pieces
    def _dec_prenet(self, frames, seed, t_ids, b_ids):
        """frames [B,L,80]; P7 dropout ALWAYS on ([TA]:284), masks keyed by (site, t, b)."""
        p = self.dec_prenet
        h = px.dropout_bits(F.relu(p.fc1(frames)), seed, px.SITE_DEC_PRENET_FC1, t_ids[None, :], b_ids[:, None])
        h = px.dropout_bits(F.relu(p.fc2(h)), seed, px.SITE_DEC_PRENET_FC2, t_ids[None, :], b_ids[:, None])
        return p.proj(h)

    def _postnet(self, mel_before, mel_lens, seed, b_ids):
        """mel_before [B,T,80] already zeroed past mel_lens -> residual [B,T,80] (P11, P9)."""
        T = mel_before.shape[1]
        m = length_mask(mel_lens, T)[:, None, :].to(torch.float32)
        t_ids = np.arange(T)
        x = mel_before.transpose(1, 2)
        n = len(self.postnet.convs)
        for i, cb in enumerate(self.postnet.convs):
            x = masked_batchnorm(cb.bn, cb.conv(x), m, self.training)
            if i < n - 1:
                x = torch.tanh(x)
            if self.training:
                x = px.dropout_bits(x.transpose(1, 2), seed, px.SITE_POSTNET_CONV0 + i,
                                    t_ids[None, :], b_ids[:, None]).transpose(1, 2)
            x = x * m
        return x.transpose(1, 2)

    # ------------------------------------------------------------------ teacher-forced
    def forward(self, phonemes, phoneme_lens, mels, mel_lens, seed: int = 0, utt_ids=None
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """phonemes [B,S] i64, phoneme_lens [B], mels [B,T,80] f32, mel_lens [B]
        -> mel_before [B,T,80], mel_after [B,T,80], stop_logits [B,T] (all zero past mel_lens)."""
        B, T, _ = mels.shape
        S = phonemes.shape[1]
        b_ids = np.arange(B) if utt_ids is None else np.asarray(utt_ids)
        t_ids = np.arange(T)
        memory = self.encode(phonemes, phoneme_lens, seed, utt_ids)
        go = torch.zeros(B, 1, mels.shape[2], dtype=mels.dtype)
        dec_in = torch.cat([go, mels[:, :-1]], dim=1)                      # P8 shift right, r = 1
        x = self._dec_prenet(dec_in, seed, t_ids, b_ids) + self.dec_alpha * self.pe[:T][None]
        x = self._drop_res(x, seed, px.SITE_DEC_PE, t_ids, b_ids)
        tvalid = length_mask(mel_lens, T)
        causal = torch.tril(torch.ones(T, T, dtype=torch.bool))
        self_mask = causal[None, None] & tvalid[:, None, None, :]          # P9 causal + key padding
        cross_mask = length_mask(phoneme_lens, S)[:, None, None, :]
        for l, layer in enumerate(self.decoder.layers):
            s0 = px.SITE_DEC_LAYER0 + 3 * l
            a = layer.self_attn(x, x, self_mask)
            x = layer.norm1(x + self._drop_res(a, seed, s0, t_ids, b_ids))
            a = layer.cross_attn(x, memory, cross_mask)
            x = layer.norm2(x + self._drop_res(a, seed, s0 + 1, t_ids, b_ids))
            f = layer.ffn(x)
            x = layer.norm3(x + self._drop_res(f, seed, s0 + 2, t_ids, b_ids))
        tm = tvalid.to(torch.float32)
        mel_before = self.mel_linear(x) * tm[..., None]
        stop_logits = self.stop_linear(x).squeeze(-1) * tm
        mel_after = (mel_before + self._postnet(mel_before, mel_lens, seed, b_ids)) * tm[..., None]
        return mel_before, mel_after, stop_logits

    # ------------------------------------------------------------------ one AR step over the KV cache
    @torch.no_grad()
    def init_decode_state(self, memory, phoneme_lens, max_len: int):
        """Hoisted cross-K/V projections + empty self-attention caches [B,H,max_len,64] per layer."""
        c = self.cfg
        B, S, _ = memory.shape
        H, dh = c.n_heads, c.d_model // c.n_heads
        layers = self.decoder.layers
        return dict(
            cross_mask=length_mask(phoneme_lens, S)[:, None, None, :],
            ck=[l.cross_attn.split(l.cross_attn.wk(memory)) for l in layers],
            cv=[l.cross_attn.split(l.cross_attn.wv(memory)) for l in layers],
            sk=[torch.zeros(B, H, max_len, dh) for _ in layers],
            sv=[torch.zeros(B, H, max_len, dh) for _ in layers])

    @torch.no_grad()
    def decode_step(self, state, frame, t: int, seed: int, b_ids):
        """frame [B,1,80] (previous mel_before frame, fp32; zeros at t = 0) -> (frame_t [B,1,80], stop logit [B])."""
        x = self._dec_prenet(frame, seed, np.array([t]), b_ids) + self.dec_alpha * self.pe[t][None, None]
        for l, layer in enumerate(self.decoder.layers):
            sa, sk, sv = layer.self_attn, state["sk"][l], state["sv"][l]
            sk[:, :, t] = sa.split(sa.wk(x))[:, :, 0]
            sv[:, :, t] = sa.split(sa.wv(x))[:, :, 0]
            a = sa.attend(sa.split(sa.wq(x)), sk[:, :, : t + 1], sv[:, :, : t + 1], None)
            x = layer.norm1(x + a)
            ca = layer.cross_attn
            a = ca.attend(ca.split(ca.wq(x)), state["ck"][l], state["cv"][l], state["cross_mask"])
            x = layer.norm2(x + a)
            x = layer.norm3(x + layer.ffn(x))
        return self.mel_linear(x), self.stop_linear(x)[:, 0, 0]                 # fp32 feedback (P8)

    # ------------------------------------------------------------------ greedy AR with KV cache
    @torch.no_grad()
    def inference(self, phonemes, phoneme_lens, max_len: int = 800, seed: int = 0, utt_ids=None,
                  return_before: bool = False):
        """-> mel_after [B,Tout,80], mel_lens [B] i32, stop_logits [B,Tout]  (P8, P10).

        Stop rule: fire when the fp32 stop logit > 0 ([TA]:849 with threshold 0.5); the firing frame
        is counted ([TA]:847: length incremented before the test); len = max_len if never; finished
        utterances keep computing but are masked; the loop ends at max_len or when all finished."""
        assert not self.training, "inference() is an eval-mode path (prenet dropout stays on regardless)"
        c = self.cfg
        B, S = phonemes.shape
        b_ids = np.arange(B) if utt_ids is None else np.asarray(utt_ids)
        memory = self.encode(phonemes, phoneme_lens, seed, utt_ids)
        state = self.init_decode_state(memory, phoneme_lens, max_len)
        frame = torch.zeros(B, 1, c.n_mels)                                     # go frame
        mel_before = torch.zeros(B, max_len, c.n_mels)
        stop_logits = torch.zeros(B, max_len)
        lens = torch.full((B,), max_len, dtype=torch.int32)
        finished = torch.zeros(B, dtype=torch.bool)
        n_steps = 0
        for t in range(max_len):
            frame, logit = self.decode_step(state, frame, t, seed, b_ids)
            mel_before[:, t] = frame[:, 0]
            stop_logits[:, t] = logit
            n_steps = t + 1
            fire = (logit > 0) & ~finished
            lens[fire] = t + 1
            finished |= fire
            if bool(finished.all()):
                break
        tm = length_mask(lens, n_steps).to(torch.float32)
        mel_before = mel_before[:, :n_steps] * tm[..., None]
        stop_logits = stop_logits[:, :n_steps] * tm
        mel_after = (mel_before + self._postnet(mel_before, lens, seed, b_ids)) * tm[..., None]
        if return_before:
            return mel_after, lens, stop_logits, mel_before
        return mel_after, lens, stop_logits


def tts_loss(mel_before, mel_after, stop_logits, mels, mel_lens, pos_weight: float = 5.0) -> torch.Tensor:
    """P13: MSE(before) + MSE(after) + BCE-with-logits(stop, pos_weight), each a mean over valid
    positions; stop target is 1 at the last valid frame only."""
    B, T, M = mels.shape
    tm = length_mask(mel_lens, T).to(torch.float32)
    n_frames = tm.sum().clamp(min=1.0)
    mse_b = (((mel_before - mels) ** 2) * tm[..., None]).sum() / (n_frames * M)
    mse_a = (((mel_after - mels) ** 2) * tm[..., None]).sum() / (n_frames * M)
    target = torch.zeros(B, T)
    target[torch.arange(B), (mel_lens.to(torch.int64) - 1).clamp(min=0)] = 1.0
    bce = F.binary_cross_entropy_with_logits(stop_logits, target, reduction="none",
                                             pos_weight=torch.tensor(pos_weight))
    return mse_b + mse_a + (bce * tm).sum() / n_frames
