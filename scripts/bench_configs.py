"""Timings of the BASELINE.json configurations that are not bench.py's headline line (those are parity-test cases,
not bench lines): teacher-forced forward (configs 0 / 3 shapes, eval arithmetic), batch-1 AR latency (config 1) and
the long-utterance stress shapes (config 4).  Usage (GPU box): python scripts/bench_configs.py -> gpurun_out/configs.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def timeit(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    model = bench.synthetic_state_dict().eval()      # (returns the module carrying the canonical synthetic weights)
    out = []
    only = os.environ.get("ONLY", "")                # substring filter on the case name (for ncu launch lists)
    for name, B, S, T in (("forward b4 s100 t400 (config 0 shape)", 4, 100, 400), ("forward b32 s100 t800 (config 3 shape, eval arithmetic)", 32, 100, 800),
                          ("forward b64 s100 t800", 64, 100, 800), ("forward b16 s300 t1600 (config 4 shape)", 16, 300, 1600)):
        if only not in name:
            continue
        g = torch.Generator().manual_seed(B * 1000 + T)
        ph = torch.randint(1, 70, (B, S), generator=g).cuda(); pl = torch.full((B,), S, dtype=torch.int32).cuda()
        mel = torch.randn(B, T, 80, generator=g).cuda(); ml = torch.full((B,), T, dtype=torch.int32).cuda()
        ms = timeit(lambda: model(ph, pl, mel, ml, seed=3), 10)
        out.append(dict(case=name, ms=ms, frames_per_s=B * T / ms * 1e3, utt_per_s=B / ms * 1e3))
        print(f"{name}: {ms:.3f} ms  {B * T / ms * 1e3:,.0f} frames/s  {B / ms * 1e3:,.1f} utt/s")
    for name, B, S, T in (("AR b1 s100 t800 (config 1, latency)", 1, 100, 800), ("AR b16 s300 t1600 (config 4)", 16, 300, 1600)):
        if only not in name:
            continue
        g = torch.Generator().manual_seed(B * 77 + T)
        ph = torch.randint(1, 70, (B, S), generator=g).cuda(); pl = torch.full((B,), S, dtype=torch.int32).cuda()
        ms = timeit(lambda: model.inference(ph, pl, max_len=T, seed=5), 3)
        out.append(dict(case=name, ms=ms, frames_per_s=B * T / ms * 1e3, us_per_decoder_step=ms * 1e3 / T))
        print(f"{name}: {ms:.2f} ms  {B * T / ms * 1e3:,.0f} frames/s  {ms * 1e3 / T:.1f} us/decoder step")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
