"""Small-shape GPU workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every hand-written kernel
of the library is launched at least once on shapes that finish in seconds under instrumentation.  No oracle, no asserts on
values beyond finiteness -- the sanitizer is the checker here (scripts/sanitize.sh drives it; logs -> profiles/).

    python scripts/sanitize_targets.py [decode|seq|train|all]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_state_dict, synthetic_inputs  # noqa: E402
from transformer_tacotron2_b200 import TransformerTTS, _lib  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
lib = _lib.load()
torch.zeros(1, device="cuda")
P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
ST = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731


def model(stop_bias=-8.0):
    src = synthetic_state_dict()
    with torch.no_grad():
        src.stop_linear.bias.fill_(stop_bias)
    m = TransformerTTS()
    m.load_state_dict(src.state_dict())
    return m


def decode():
    m = model()
    for B, S, T, G in [(1, 10, 20, 0), (5, 33, 70, 0), (7, 20, 40, 3), (9, 24, 30, 4), (3, 17, 130, 1), (70, 12, 18, 0)]:
        m.set_option("cluster_group", G)
        ph, pl = synthetic_inputs(B, S, 11 + B)
        pl = torch.randint(max(1, S // 2), S + 1, (B,), dtype=torch.int32); pl[0] = S
        out = m.inference(ph.cuda(), pl.cuda(), max_len=T, seed=3)
        assert torch.isfinite(out[0]).all()
        out = m.inference(ph, pl, max_len=T, seed=3)                      # host path
        assert torch.isfinite(out[0]).all()
        print("decode", B, S, T, G, "ok", flush=True)
    # early stop inside groups + resume in chunks through the C ABI
    m2 = model(stop_bias=-0.45)
    ph, pl = synthetic_inputs(6, 20, 5)
    out = m2.inference(ph.cuda(), pl.cuda(), max_len=60, seed=9)
    print("decode stopping lens", out[1].tolist(), flush=True)
    l = m._ensure_handle()
    B, S, T = 4, 16, 48
    ph, pl = synthetic_inputs(B, S, 21)
    ws = m._workspace(B, S, T)
    m._check(l.tts_decode_begin(m._handle, ws.data_ptr(), B, S, T, 3, 0, ST()), "begin")
    m._check(l.tts_encode(m._handle, ws.data_ptr(), P(ph.cuda()), P(pl.cuda()), B, S, T, None, ST()), "encode")
    for _ in range(0, T, 7):
        m._check(l.tts_decode_steps(m._handle, ws.data_ptr(), 7, ST()), "steps")
    torch.cuda.synchronize()
    print("decode resume ok", flush=True)


def seq():
    g = torch.Generator().manual_seed(1)
    for M, N, K in [(130, 512, 512), (257, 384, 72), (33, 896, 8)]:
        A = (torch.randn(M, K, generator=g) * .5).to(torch.bfloat16).cuda(); W = (torch.randn(N, K, generator=g) * .05).to(torch.bfloat16).cuda()
        bias = torch.randn(N, generator=g).cuda(); out = torch.empty(M, N, device="cuda")
        assert lib.tts_k_gemm(P(A), P(W), P(bias), P(out), M, N, K, 1, ST()) == 0
    for B, T, Cin, Cout in [(2, 37, 96, 512), (3, 130, 512, 128)]:
        X = (torch.randn(B, T, Cin, generator=g) * .5).to(torch.bfloat16).cuda(); W = (torch.randn(5, Cout, Cin, generator=g) * .03).to(torch.bfloat16).cuda()
        bias = torch.randn(Cout, generator=g).cuda(); lens = torch.full((B,), T, dtype=torch.int32).cuda(); Y = torch.empty(B, T, Cout, device="cuda")
        assert lib.tts_k_conv5(P(X), P(W), P(bias), P(lens), P(Y), B, T, Cin, Cout, 2, ST()) == 0
    for B, Lq, Lk, causal in [(2, 100, 100, 0), (2, 129, 129, 1), (1, 64, 300, 0), (1, 260, 260, 1)]:
        H = 8
        q, k, v, do = ((torch.randn(B, L, H * 64, generator=g) * .5).to(torch.bfloat16).cuda() for L in (Lq, Lk, Lk, Lq))
        o = torch.empty_like(q); lse = torch.empty(B, H, Lq, device="cuda")
        kl = torch.randint(1, Lk + 1, (B,), generator=g, dtype=torch.int32).cuda()
        assert lib.tts_k_attention_lse(P(q), P(k), P(v), P(o), P(lse), P(kl), B, H, Lq, Lk, causal, ST()) == 0
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        scratch = torch.empty(B * Lq * H * 64 + B * H * Lq + 64, device="cuda")
        assert lib.tts_k_attention_bwd(P(q), P(k), P(v), P(o), P(do), P(lse), P(kl), P(dq), P(dk), P(dv), P(scratch), B, H, Lq, Lk, causal, ST()) == 0
    x = torch.randn(77, 512, generator=g).cuda(); y = torch.empty(77, 512, dtype=torch.bfloat16, device="cuda")
    assert lib.tts_k_layernorm(P(x), P(torch.ones(512).cuda()), P(torch.zeros(512).cuda()), P(y), 77, 1e-5, ST()) == 0
    torch.cuda.synchronize()
    m = model()
    ph, pl = synthetic_inputs(3, 37, 4)
    mels = torch.randn(3, 129, 80, generator=g); ml = torch.tensor([129, 77, 100], dtype=torch.int32)
    out = m(ph, pl, mels, ml, seed=5)
    assert torch.isfinite(out[1]).all()
    print("seq ok", flush=True)


def train():
    from transformer_tacotron2_b200.training import Trainer
    m = model()
    tr = Trainer(m, lr=1e-4)
    g = torch.Generator().manual_seed(2)
    B, S, T = 3, 20, 70
    ph, pl = synthetic_inputs(B, S, 8)
    mels = torch.randn(B, T, 80, generator=g); ml = torch.tensor([70, 33, 51], dtype=torch.int32)
    m.set_option("train_graph", 0)
    for i in range(2):
        loss = tr.step(ph, pl, mels, ml, seed=i)
    print("train ok", float(loss), flush=True)


if what in ("decode", "all"):
    decode()
if what in ("seq", "all"):
    seq()
if what in ("train", "all"):
    train()
torch.cuda.synchronize()
print("sanitize_targets done", flush=True)
