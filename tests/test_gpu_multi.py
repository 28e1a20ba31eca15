"""Multi-GPU correctness through real processes (skipped on boxes with fewer than 2 GPUs): one process per GPU, NCCL rendezvous on
127.0.0.1.  (a) the fused peer-memory optimiser step (reduce-scatter -> Adam -> all-gather over NVLink, CUDA IPC) leaves every
rank with the parameters of the NCCL all-reduce + Adam path, and every rank equal to rank 0 (scripts/dp_peer_check.py asserts both);
(b) bench.py at N = 2 exits 0 and prints one JSON line whose inference value is ~2x... (weak scaling, no collective)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need2():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")


def _torchrun(args, port, timeout):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + args
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def test_peer_adam_equals_nccl_allreduce_adam_two_gpus():
    _need2()
    r = _torchrun(["scripts/dp_peer_check.py"], 29541, 600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dp_peer_check ok" in r.stdout
    line = [l for l in r.stdout.splitlines() if "max |fused - nccl|" in l][-1]
    print(line)


def test_bench_two_gpus_exits_zero():
    _need2()
    r = _torchrun(["bench.py", "--gpus", "2", "--steps", "2", "--warmup", "3", "--no-extra"], 29542, 900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["n_gpus"] == 2 and d["value"] > 0 and d["train"]["value"] > 0
