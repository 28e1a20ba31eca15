"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/tts_b200.h
declares (no compute calls without a GPU)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "tts_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tts_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    from transformer_tacotron2_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, f"{n} declared in the header but not bound in _lib.SIGNATURES"
    assert set(_lib.SIGNATURES) == set(names)
    assert b"sm_100a" in lib.tts_version()


def test_library_contains_sm100a_code():
    import subprocess
    from transformer_tacotron2_b200 import _lib
    _lib.load()
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_gpu_means_loud_failure_not_fallback():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from transformer_tacotron2_b200 import TransformerTTS
    m = TransformerTTS()
    with pytest.raises(RuntimeError):
        m.inference(torch.zeros(1, 4, dtype=torch.int64), torch.tensor([4]), max_len=2)
    # tts_create itself refuses without an sm_100 device
    from transformer_tacotron2_b200 import _lib
    lib = _lib.load()
    cc = _lib.TtsConfig(C.sizeof(_lib.TtsConfig), 128, 512, 8, 6, 6, 2048, 80, 256, 3, 5, 512, 5, 2048, 1e-5, 1e-5)
    h = C.c_void_p()
    assert lib.tts_create(C.byref(cc), 0, C.byref(h)) != 0


def test_training_without_gpu_fails_loudly_too():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from transformer_tacotron2_b200 import TransformerTTS
    from transformer_tacotron2_b200.training import Trainer
    with pytest.raises(RuntimeError):
        Trainer(TransformerTTS())
    # the training entry points refuse a null handle instead of crashing
    from transformer_tacotron2_b200 import _lib
    lib = _lib.load()
    assert lib.tts_train_begin(None) != 0
    assert lib.tts_train_num_tensors(None) == -1
    assert lib.tts_train_workspace_bytes(None, 4, 10, 20) == 0


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "transformer_tacotron2_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_kv_cache_block_layout():
    """The decode kernel's cache block layout (host arithmetic of the shipped library, no GPU): a bijection onto 8192 elements;
    rows [0, 16 n) of K and V together fill exactly the first 16 n * 128 elements (so any ring stage is ONE contiguous copy);
    K rows are row-major inside a 16-row sub-chunk; the V half of a sub-chunk is in mma.m16n8k16 A-fragment order of V^T."""
    from transformer_tacotron2_b200 import _lib
    lib = _lib.load()
    idx = {(w, r, d): lib.tts_debug_kv_index(r, d, w) for w in (0, 1) for r in range(64) for d in range(64)}
    assert sorted(idx.values()) == list(range(8192))
    assert lib.tts_debug_kv_index(64, 0, 0) == -1 and lib.tts_debug_kv_index(0, 64, 1) == -1 and lib.tts_debug_kv_index(0, 0, 2) == -1
    for n in range(1, 5):
        got = sorted(v for (w, r, d), v in idx.items() if r < 16 * n)
        assert got == list(range(16 * n * 128))
    for r in range(64):
        base = (r // 16) * 2048 + (r % 16) * 64
        assert [idx[(0, r, d)] for d in range(64)] == list(range(base, base + 64))
    for s in range(4):                     # 16-row sub-chunk
        for dt in range(4):                # 16-dim tile of V^T
            for g in range(8):
                for t4 in range(4):
                    base = s * 2048 + 1024 + (dt * 32 + g * 4 + t4) * 8
                    want = [(16 * s + 4 * t4 + 0, 16 * dt + g), (16 * s + 4 * t4 + 1, 16 * dt + g),            # a0: A[g][2 t4, 2 t4 + 1]
                            (16 * s + 4 * t4 + 0, 16 * dt + g + 8), (16 * s + 4 * t4 + 1, 16 * dt + g + 8),    # a1: A[g + 8][..]
                            (16 * s + 4 * t4 + 2, 16 * dt + g), (16 * s + 4 * t4 + 3, 16 * dt + g),            # a2: A[g][2 t4 + 8, + 9]
                            (16 * s + 4 * t4 + 2, 16 * dt + g + 8), (16 * s + 4 * t4 + 3, 16 * dt + g + 8)]    # a3
                    assert [idx[(1, r, d)] for r, d in want] == list(range(base, base + 8))
