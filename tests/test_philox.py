"""T2: RNG contract known-answer tests (Random123 kat_vectors, philox4x32 10 rounds)."""
import numpy as np
import torch

from oracle import philox as px

KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
     [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


def test_random123_known_answers():
    for ctr, key, want in KAT:
        got = px.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert [int(v) for v in got] == want


def test_vectorised_equals_scalar():
    rng = np.random.default_rng(0)
    ctr = rng.integers(0, 2**32, size=(5, 7, 4), dtype=np.uint64).astype(np.uint32)
    key = np.array([123, 456], dtype=np.uint32)
    got = px.philox4x32_10(ctr, key)
    for i in range(5):
        for j in range(7):
            assert (px.philox4x32_10(ctr[i, j], key) == got[i, j]).all()


def test_bit_site_layout():
    """channel c <-> chunk c//128, word (c%128)//32, bit c%32; counter = (site, t, b, chunk)."""
    seed, site, t, b = 7 + (5 << 32), px.SITE_DEC_PRENET_FC2, 13, 3
    m = px.keep_mask_bits(seed, site, np.array([t]), np.array([b]), 256)[0]
    key = np.array([7, 5], dtype=np.uint32)
    for c in (0, 1, 31, 32, 127, 128, 200, 255):
        w = px.philox4x32_10(np.array([site, t, b, c // 128], dtype=np.uint32), key)
        assert int(m[c]) == (int(w[(c % 128) // 32]) >> (c % 32)) & 1
    assert 0.35 < float(m.mean()) < 0.65


def test_word_site_layout_and_rate():
    seed, site = 99, px.SITE_DEC_LAYER0 + 4
    t, b = np.arange(50), np.arange(3)
    m = px.keep_mask_words(seed, site, t[None, :], b[:, None], 512, 0.1)
    assert m.shape == (3, 50, 512)
    assert abs(float(m.mean()) - 0.9) < 0.01
    key = np.array([99, 0], dtype=np.uint32)
    w = px.philox4x32_10(np.array([site, 17, 2, 100], dtype=np.uint32), key)
    for j in range(4):
        assert int(m[2, 17, 400 + j]) == int(int(w[j]) >= int(0.1 * 2**32))


def test_masks_depend_on_global_utterance_id_only():
    a = px.keep_mask_bits(7, 0, np.arange(4)[None, :], np.array([5, 6])[:, None], 256)
    b = px.keep_mask_bits(7, 0, np.arange(4)[None, :], np.array([6])[:, None], 256)
    assert torch.equal(a[1], b[0])
