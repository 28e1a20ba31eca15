"""T3: per-kernel parity on a real B200 (pytest -m gpu).  Each CUDA kernel is called through the
C ABI (tts_k_*) and compared with a plain PyTorch fp32 computation of the same op on the same
bf16-rounded inputs; integer work (Philox keep-bits) must be bit-exact against the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from transformer_tacotron2_b200 import _lib
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    return _lib.load()


def _p(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("M,N,K,act", [(128, 128, 32, 0), (300, 512, 512, 1), (1, 128, 96, 0), (777, 1536, 512, 0),
                                         (130, 512, 2048, 2), (6400, 256, 256, 1), (257, 384, 72, 0), (4097, 640, 520, 1),
                                         (33, 896, 8, 0)])
def test_gemm(lib, M, N, K, act):
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    W = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).cuda()
    bias = torch.randn(N, generator=g).cuda()
    Cout = torch.empty(M, N, device="cuda")
    rc = lib.tts_k_gemm(_p(A), _p(W), _p(bias), _p(Cout), M, N, K, act, _stream())
    assert rc == 0
    ref = A.float() @ W.float().T + bias
    ref = torch.relu(ref) if act == 1 else torch.tanh(ref) if act == 2 else ref
    torch.cuda.synchronize()
    assert torch.allclose(Cout, ref, atol=2e-3, rtol=2e-3), float((Cout - ref).abs().max())


@pytest.mark.parametrize("B,T,Cin,Cout", [(2, 37, 96, 512), (3, 130, 512, 512), (1, 5, 512, 128), (4, 64, 512, 128), (2, 129, 80, 128),
                                          (5, 300, 512, 256), (3, 1, 512, 384)])
def test_conv5(lib, B, T, Cin, Cout):
    g = torch.Generator().manual_seed(B * T)
    X = (torch.randn(B, T, Cin, generator=g) * 0.5).to(torch.bfloat16)
    Wt = (torch.randn(Cout, Cin, 5, generator=g) * 0.03).to(torch.bfloat16)
    bias = torch.randn(Cout, generator=g)
    lens = torch.randint(1, T + 1, (B,), generator=g, dtype=torch.int32); lens[0] = T
    m = (torch.arange(T)[None, :] < lens[:, None]).float()
    Xm = X.float() * m[..., None]                          # activations are zero past the length (P9)
    ref = torch.nn.functional.conv1d(Xm.transpose(1, 2), Wt.float(), bias, padding=2).transpose(1, 2)
    ref = torch.tanh(ref) * m[..., None]
    Wp = Wt.permute(2, 0, 1).contiguous().cuda()           # [5][Cout][Cin]
    Xd = Xm.to(torch.bfloat16).cuda(); Y = torch.empty(B, T, Cout, device="cuda")
    bias_d, lens_d = bias.cuda(), lens.cuda()               # keep every device buffer alive across the call
    rc = lib.tts_k_conv5(_p(Xd), _p(Wp), _p(bias_d), _p(lens_d), _p(Y), B, T, Cin, Cout, 2, _stream())
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.allclose(Y.cpu(), ref, atol=3e-3, rtol=3e-3), float((Y.cpu() - ref).abs().max())


@pytest.mark.parametrize("B,Lq,Lk,causal", [(2, 100, 100, 0), (1, 1, 1, 0), (2, 63, 65, 0), (2, 129, 129, 1),
                                             (1, 400, 400, 1), (2, 200, 37, 0), (1, 64, 300, 0), (1, 127, 127, 1), (2, 128, 128, 0),
                                             (1, 257, 129, 0), (1, 1600, 1600, 1), (3, 800, 100, 0)])
def test_attention(lib, B, Lq, Lk, causal):
    H = 8
    g = torch.Generator().manual_seed(Lq * 7 + Lk)
    Q = torch.randn(B, Lq, H * 64, generator=g).to(torch.bfloat16)
    K = torch.randn(B, Lk, H * 64, generator=g).to(torch.bfloat16)
    V = torch.randn(B, Lk, H * 64, generator=g).to(torch.bfloat16)
    klens = torch.randint(1, Lk + 1, (B,), generator=g, dtype=torch.int32); klens[0] = Lk
    q = Q.float().view(B, Lq, H, 64).transpose(1, 2); k = K.float().view(B, Lk, H, 64).transpose(1, 2)
    v = V.float().view(B, Lk, H, 64).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / 8.0
    mask = (torch.arange(Lk)[None, :] < klens[:, None])[:, None, None, :]
    if causal:
        mask = mask & torch.tril(torch.ones(Lq, Lk, dtype=torch.bool))[None, None]
    s = s.masked_fill(~mask, float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, Lq, H * 64)
    O = torch.empty(B, Lq, H * 64, dtype=torch.bfloat16, device="cuda")
    Qd, Kd, Vd, kl = Q.cuda(), K.cuda(), V.cuda(), klens.cuda()
    rc = lib.tts_k_attention(_p(Qd), _p(Kd), _p(Vd), _p(O), _p(kl), B, H, Lq, Lk, causal, _stream())
    assert rc == 0
    torch.cuda.synchronize()
    got = O.float().cpu()
    assert torch.allclose(got, ref, atol=2e-2, rtol=2e-2), float((got - ref).abs().max())


@pytest.mark.parametrize("B,Lq,Lk,causal", [(2, 100, 100, 0), (2, 129, 129, 1), (1, 400, 400, 1), (2, 200, 37, 0), (1, 64, 300, 0),
                                             (2, 260, 260, 1), (1, 1, 1, 0), (1, 127, 127, 1), (2, 128, 128, 0), (1, 257, 129, 0),
                                             (1, 800, 800, 1), (3, 800, 100, 0)])
def test_attention_backward(lib, B, Lq, Lk, causal):
    """dQ / dK / dV of the tcgen05 backward kernel against torch autograd (fp32) on the same bf16 inputs."""
    H = 8
    g = torch.Generator().manual_seed(Lq * 11 + Lk)
    Q = torch.randn(B, Lq, H * 64, generator=g).to(torch.bfloat16)
    K = torch.randn(B, Lk, H * 64, generator=g).to(torch.bfloat16)
    V = torch.randn(B, Lk, H * 64, generator=g).to(torch.bfloat16)
    dO = torch.randn(B, Lq, H * 64, generator=g).to(torch.bfloat16)
    klens = torch.randint(1, Lk + 1, (B,), generator=g, dtype=torch.int32); klens[0] = Lk
    qf, kf, vf = (x.float().requires_grad_(True) for x in (Q, K, V))
    q = qf.view(B, Lq, H, 64).transpose(1, 2); k = kf.view(B, Lk, H, 64).transpose(1, 2); v = vf.view(B, Lk, H, 64).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / 8.0
    mask = (torch.arange(Lk)[None, :] < klens[:, None])[:, None, None, :]
    if causal:
        mask = mask & torch.tril(torch.ones(Lq, Lk, dtype=torch.bool))[None, None]
    s = s.masked_fill(~mask, float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, Lq, H * 64)
    ref.backward(dO.float())
    Qd, Kd, Vd, dOd, kl = Q.cuda(), K.cuda(), V.cuda(), dO.cuda(), klens.cuda()
    O = torch.empty_like(Qd); lse = torch.empty(B, H, Lq, device="cuda")
    assert lib.tts_k_attention_lse(_p(Qd), _p(Kd), _p(Vd), _p(O), _p(lse), _p(kl), B, H, Lq, Lk, causal, _stream()) == 0
    dQ, dK, dV = torch.empty_like(Qd), torch.empty_like(Kd), torch.empty_like(Vd)
    scratch = torch.empty(B * Lq * H * 64 + B * H * Lq, device="cuda")
    rc = lib.tts_k_attention_bwd(_p(Qd), _p(Kd), _p(Vd), _p(O), _p(dOd), _p(lse), _p(kl), _p(dQ), _p(dK), _p(dV), _p(scratch),
                                 B, H, Lq, Lk, causal, _stream())
    assert rc == 0
    torch.cuda.synchronize()
    ref_lse = torch.logsumexp(s, -1) * 1.4426950408889634
    assert torch.allclose(lse.cpu(), ref_lse, atol=2e-2, rtol=1e-3)
    for name, got, want in (("dQ", dQ, qf.grad), ("dK", dK, kf.grad), ("dV", dV, vf.grad)):
        got = got.float().cpu()
        if float(want.norm()) < 1e-6:                         # a single key: softmax = 1, dS = 0, so dQ = dK = 0 exactly
            assert float(got.norm()) < 1e-3, (name, float(got.norm()))
            continue
        err = float((got - want).norm() / want.norm())
        assert err < 2e-2, (name, err)                       # bf16 P / dS operands: ~1e-2 relative
        assert torch.allclose(got, want, atol=6e-2, rtol=6e-2), (name, float((got - want).abs().max()))


def test_layernorm(lib):
    g = torch.Generator().manual_seed(3)
    X = torch.randn(333, 512, generator=g) * 3 + 1
    gam, bet = torch.randn(512, generator=g), torch.randn(512, generator=g)
    Y = torch.empty(333, 512, dtype=torch.bfloat16, device="cuda")
    Xd, gd, bd = X.cuda(), gam.cuda(), bet.cuda()
    rc = lib.tts_k_layernorm(_p(Xd), _p(gd), _p(bd), _p(Y), 333, 1e-5, _stream())
    assert rc == 0
    ref = torch.nn.functional.layer_norm(X, (512,), gam, bet, 1e-5)
    torch.cuda.synchronize()
    assert torch.allclose(Y.float().cpu(), ref, atol=3e-2, rtol=1e-2)


def test_philox_bits_bit_exact_vs_oracle(lib):
    from oracle import philox as px
    T, B, Cn, seed, site, off = 9, 5, 256, 7 + (3 << 32), px.SITE_DEC_PRENET_FC2, 11
    out = torch.empty(T, B, Cn, dtype=torch.uint8, device="cuda")
    rc = lib.tts_k_philox_bits(seed, site, T, B, Cn, off, _p(out), _stream())
    assert rc == 0
    want = px.keep_mask_bits(seed, site, np.arange(T)[:, None], (off + np.arange(B))[None, :], Cn)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu().float(), want)


def test_argument_errors(lib):
    assert lib.tts_k_gemm(None, None, None, None, 1, 1, 1, 0, None) < 0
    h = C.c_void_p()
    assert lib.tts_create(None, 0, C.byref(h)) < 0
