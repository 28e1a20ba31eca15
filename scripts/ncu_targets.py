"""One launch (after two warm-ups) of each sequence-parallel kernel at its headline shape, for `ncu --set full` captures.
Usage (GPU box): ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|flash_attn" -o out python scripts/ncu_targets.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_tacotron2_b200 import _lib  # noqa: E402

lib = _lib.load()
torch.zeros(1, device="cuda")
P = lambda t: C.c_void_p(t.data_ptr())
ST = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
REPS = 3


def gemm(M, N, K):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16); W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    bias = torch.zeros(N, device="cuda"); Cc = torch.empty(M, N, device="cuda")
    for _ in range(REPS):
        assert lib.tts_k_gemm(P(A), P(W), P(bias), P(Cc), M, N, K, 0, ST()) == 0
    torch.cuda.synchronize()


def conv(B, T, Cin, Cout):
    X = torch.randn(B, T, Cin, device="cuda").to(torch.bfloat16); W = (torch.randn(5, Cout, Cin, device="cuda") * 0.02).to(torch.bfloat16)
    bias = torch.zeros(Cout, device="cuda"); lens = torch.full((B,), T, dtype=torch.int32, device="cuda"); Y = torch.empty(B, T, Cout, device="cuda")
    for _ in range(REPS):
        assert lib.tts_k_conv5(P(X), P(W), P(bias), P(lens), P(Y), B, T, Cin, Cout, 2, ST()) == 0
    torch.cuda.synchronize()


def attn(B, L, causal, bwd=False):
    H = 8
    Q = torch.randn(B, L, H * 64, device="cuda").to(torch.bfloat16); K_ = torch.randn_like(Q); V = torch.randn_like(Q); O = torch.empty_like(Q)
    kl = torch.full((B,), L, dtype=torch.int32, device="cuda"); lse = torch.empty(B, H, L, device="cuda")
    for _ in range(REPS):
        assert lib.tts_k_attention_lse(P(Q), P(K_), P(V), P(O), P(lse), P(kl), B, H, L, L, causal, ST()) == 0
    if bwd:
        dO = torch.randn_like(Q); dQ, dK, dV = torch.empty_like(Q), torch.empty_like(Q), torch.empty_like(Q)
        scratch = torch.empty(B * L * H * 64 + B * H * L, device="cuda")
        for _ in range(REPS):
            assert lib.tts_k_attention_bwd(P(Q), P(K_), P(V), P(O), P(dO), P(lse), P(kl), P(dQ), P(dK), P(dV), P(scratch), B, H, L, L, causal, ST()) == 0
    torch.cuda.synchronize()


gemm(25600, 1536, 512)
gemm(25600, 512, 2048)
conv(64, 800, 512, 512)
attn(8, 4096, 0)
attn(64, 800, 1, bwd=True)
print("done")
