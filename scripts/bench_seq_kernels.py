"""Throughput of the sequence-parallel kernels (tcgen05 GEMM / conv, flash attention) at the BASELINE shapes.
Usage (GPU box): python scripts/bench_seq_kernels.py  -> table + gpurun_out/seq_kernels.json"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformer_tacotron2_b200 import _lib  # noqa: E402

lib = _lib.load()
torch.zeros(1, device="cuda")
P = lambda t: C.c_void_p(t.data_ptr())
ST = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
PEAK = 1418.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
except Exception:
    pass


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = []
for name, fn_name in (("tcgen05", "tts_k_gemm"),):
    if not hasattr(lib, fn_name):
        continue
    for (M, N, K) in [(25600, 1536, 512), (25600, 2048, 512), (25600, 512, 2048), (51200, 512, 512), (6400, 6144, 512)]:
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16); W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        bias = torch.zeros(N, device="cuda"); Cc = torch.empty(M, N, device="cuda")
        ms = timeit(lambda: getattr(lib, fn_name)(P(A), P(W), P(bias), P(Cc), M, N, K, 0, ST()))
        tf = 2.0 * M * N * K / ms / 1e9
        out.append(dict(kernel=f"gemm {name}", M=M, N=N, K=K, ms=ms, tflops=tf, frac_of_measured_peak=tf / PEAK))
        print(f"gemm {name:9s} M={M:6d} N={N:5d} K={K:5d}: {ms:7.3f} ms  {tf:7.1f} TFLOP/s  ({tf / PEAK:.1%} of {PEAK:.0f})")
for name, fn_name in (("tcgen05", "tts_k_conv5"),):
    if not hasattr(lib, fn_name):
        continue
    B, T, Cin, Cout = 64, 800, 512, 512
    X = torch.randn(B, T, Cin, device="cuda").to(torch.bfloat16); W = (torch.randn(5, Cout, Cin, device="cuda") * 0.02).to(torch.bfloat16)
    bias = torch.zeros(Cout, device="cuda"); lens = torch.full((B,), T, dtype=torch.int32, device="cuda"); Y = torch.empty(B, T, Cout, device="cuda")
    ms = timeit(lambda: getattr(lib, fn_name)(P(X), P(W), P(bias), P(lens), P(Y), B, T, Cin, Cout, 2, ST()))
    tf = 2.0 * B * T * Cout * Cin * 5 / ms / 1e9
    out.append(dict(kernel=f"conv5 {name}", B=B, T=T, Cin=Cin, Cout=Cout, ms=ms, tflops=tf, frac_of_measured_peak=tf / PEAK))
    print(f"conv5 {name:9s} B={B} T={T} {Cin}->{Cout}: {ms:7.3f} ms  {tf:7.1f} TFLOP/s  ({tf / PEAK:.1%})")
for (B, L, causal) in [(32, 400, 1), (32, 800, 1), (16, 1600, 1), (64, 100, 0), (64, 800, 1), (32, 800, 0), (8, 4096, 0)]:
    H = 8
    Q = torch.randn(B, L, H * 64, device="cuda").to(torch.bfloat16); K_ = torch.randn_like(Q); V = torch.randn_like(Q); O = torch.empty_like(Q)
    kl = torch.full((B,), L, dtype=torch.int32, device="cuda")
    ms = timeit(lambda: lib.tts_k_attention(P(Q), P(K_), P(V), P(O), P(kl), B, H, L, L, causal, ST()))
    fl = 4.0 * B * H * 64 * (L * (L + 1) / 2 if causal else L * L)
    tf = fl / ms / 1e9
    out.append(dict(kernel="flash attention fwd", B=B, L=L, causal=causal, ms=ms, tflops=tf, frac_of_measured_peak=tf / PEAK))
    print(f"attention B={B} L={L} causal={causal}: {ms:7.3f} ms  {tf:7.1f} TFLOP/s  ({tf / PEAK:.1%})")
for (B, L, causal) in [(32, 800, 1), (64, 800, 1), (8, 4096, 0)]:
    H = 8
    Q = torch.randn(B, L, H * 64, device="cuda").to(torch.bfloat16); K_ = torch.randn_like(Q); V = torch.randn_like(Q); O = torch.empty_like(Q)
    dO = torch.randn_like(Q); dQ, dK, dV = torch.empty_like(Q), torch.empty_like(Q), torch.empty_like(Q)
    kl = torch.full((B,), L, dtype=torch.int32, device="cuda"); lse = torch.empty(B, H, L, device="cuda")
    scratch = torch.empty(B * L * H * 64 + B * H * L, device="cuda")
    lib.tts_k_attention_lse(P(Q), P(K_), P(V), P(O), P(lse), P(kl), B, H, L, L, causal, ST())
    ms = timeit(lambda: lib.tts_k_attention_bwd(P(Q), P(K_), P(V), P(O), P(dO), P(lse), P(kl), P(dQ), P(dK), P(dV), P(scratch), B, H, L, L, causal, ST()))
    fl = 10.0 * B * H * 64 * (L * (L + 1) / 2 if causal else L * L)
    tf = fl / ms / 1e9
    out.append(dict(kernel="flash attention bwd (incl. dsum / dQ cast passes)", B=B, L=L, causal=causal, ms=ms, tflops=tf, frac_of_measured_peak=tf / PEAK))
    print(f"attention bwd B={B} L={L} causal={causal}: {ms:7.3f} ms  {tf:7.1f} TFLOP/s  ({tf / PEAK:.1%})")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/seq_kernels.json", "w"), indent=1)
