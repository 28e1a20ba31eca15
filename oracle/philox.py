"""ORACLE (test infrastructure, not product code) -- counter-based dropout RNG contract.

Spec source: SURVEY.md section 8-P, decision P16 (the reference repository ships no code:
/root/reference/README.md:1-3 is its entire content, so there is no reference RNG to follow).

Contract (identical in the CUDA kernels, see transformer_tacotron2_b200/csrc/philox.cuh):

* generator  : Philox4x32-10 (Salmon et al., SC'11; Random123 `philox4x32_R(10, ...)`).
* key        : (seed & 0xffffffff, seed >> 32).
* counter    : (site_id, t, b, chunk) -- site_id names the dropout site (SITE_* below), t is the
               time index of the row (decoder frame index / phoneme position), b is the GLOBAL
               utterance id (so that a sharded batch draws the same bits as the unsharded one),
               chunk selects a group of channels.
* p = 0.5 sites ("bit sites"): one Philox call yields 4 x 32 = 128 keep-bits; channel
               c lives in chunk c // 128, word (c % 128) // 32, bit c % 32; keep iff the bit is 1;
               kept values are scaled by 2.0.
* p = 0.1 sites ("word sites"): one Philox call yields 4 words for 4 consecutive channels
               (chunk = c // 4, word = c % 4); DROP iff word < floor(p * 2**32); kept values are
               scaled by 1 / (1 - p) (computed in fp32).

Pinned by the Random123 known-answer vectors in tests/test_philox.py ("parity unpinned" by the
reference itself, which has no tests).
"""
from __future__ import annotations

import numpy as np
import torch

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)
_SHIFT32 = np.uint64(32)

# ---- dropout site ids (shared with csrc/philox.cuh) -------------------------------------------
SITE_DEC_PRENET_FC1 = 0      # p = 0.5, ALWAYS on (training and inference)
SITE_DEC_PRENET_FC2 = 1      # p = 0.5, ALWAYS on
SITE_ENC_PRENET_CONV0 = 2    # +i for conv i in 0..2, p = 0.5, training only
SITE_POSTNET_CONV0 = 5       # +i for conv i in 0..4, p = 0.5, training only
SITE_ENC_PE = 16             # p = 0.1, training only
SITE_DEC_PE = 17             # p = 0.1, training only
SITE_ENC_LAYER0 = 32         # + 2*layer + {0: self-attn, 1: ffn}, p = 0.1, training only
SITE_DEC_LAYER0 = 64         # + 3*layer + {0: self-attn, 1: cross-attn, 2: ffn}, p = 0.1, training only


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Vectorised Philox4x32-10. counter [...,4] uint32, key [...,2] uint32 (broadcastable)
    -> [...,4] uint32."""
    c = np.asarray(counter, dtype=np.uint64)
    k = np.asarray(key, dtype=np.uint64)
    c0, c1, c2, c3 = (c[..., i].copy() for i in range(4))
    k0 = np.broadcast_to(k[..., 0], c0.shape).copy()
    k1 = np.broadcast_to(k[..., 1], c0.shape).copy()
    for r in range(10):
        if r > 0:
            k0 = (k0 + np.uint64(PHILOX_W0)) & _MASK32
            k1 = (k1 + np.uint64(PHILOX_W1)) & _MASK32
        p0 = PHILOX_M0 * c0          # 32x32 -> 64 bit products, exact in uint64
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> _SHIFT32, p0 & _MASK32
        hi1, lo1 = p1 >> _SHIFT32, p1 & _MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK32, lo1, (hi0 ^ c3 ^ k1) & _MASK32, lo0
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def _key(seed: int) -> np.ndarray:
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint32)


def _words(seed: int, site: int, t: np.ndarray, b: np.ndarray, n_chunks: int) -> np.ndarray:
    """Philox output words for rows (t, b): returns uint32 [rows..., n_chunks * 4]."""
    t = np.asarray(t, dtype=np.uint32)
    b = np.asarray(b, dtype=np.uint32)
    shape = np.broadcast_shapes(t.shape, b.shape)
    ctr = np.empty(shape + (n_chunks, 4), dtype=np.uint32)
    ctr[..., 0] = np.uint32(site)
    ctr[..., 1] = np.broadcast_to(t, shape)[..., None]
    ctr[..., 2] = np.broadcast_to(b, shape)[..., None]
    ctr[..., 3] = np.arange(n_chunks, dtype=np.uint32)
    out = philox4x32_10(ctr, _key(seed))
    return out.reshape(shape + (n_chunks * 4,))


def keep_mask_bits(seed: int, site: int, t, b, n_channels: int) -> torch.Tensor:
    """p = 0.5 site: float32 keep mask (0/1) of shape broadcast(t, b) + [n_channels]."""
    n_chunks = (n_channels + 127) // 128
    w = _words(seed, site, t, b, n_chunks)                      # [..., n_chunks*4]
    bit = np.arange(32, dtype=np.uint32)
    bits = (w[..., None] >> bit) & np.uint32(1)                 # [..., words, 32]
    bits = bits.reshape(w.shape[:-1] + (n_chunks * 128,))[..., :n_channels]
    return torch.from_numpy(bits.astype(np.float32))


def keep_mask_words(seed: int, site: int, t, b, n_channels: int, p: float) -> torch.Tensor:
    """p-site with one u32 per channel: float32 keep mask (0/1)."""
    assert n_channels % 4 == 0
    thresh = np.uint32(int(p * 4294967296.0))
    w = _words(seed, site, t, b, n_channels // 4)               # [..., n_channels]
    return torch.from_numpy((w >= thresh).astype(np.float32))


def dropout_bits(x: torch.Tensor, seed: int, site: int, t, b) -> torch.Tensor:
    """Inverted dropout, p = 0.5, on x[..., C]; t and b broadcast against x.shape[:-1]."""
    m = keep_mask_bits(seed, site, t, b, x.shape[-1])
    return x * m.to(x.dtype) * 2.0


def dropout_words(x: torch.Tensor, seed: int, site: int, t, b, p: float) -> torch.Tensor:
    m = keep_mask_words(seed, site, t, b, x.shape[-1], p)
    scale = torch.tensor(1.0, dtype=torch.float32) / torch.tensor(1.0 - p, dtype=torch.float32)
    return x * m.to(x.dtype) * scale.to(x.dtype)
