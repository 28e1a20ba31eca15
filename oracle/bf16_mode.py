"""TEST INFRASTRUCTURE (see oracle/__init__.py): an operand-rounding mode for the CPU oracle.

The B200 path feeds every tensor-core product with bf16 operands and accumulates in fp32; the oracle computes in
fp32 throughout (BASELINE.json north_star).  Inside `with bf16_operands():` every matrix product of the oracle -- F.linear,
F.conv1d, torch.matmul (attention scores and the probability-value product) -- sees both operands rounded to bf16 and still
accumulates in fp32.  It is NOT the parity baseline: the tests use it to show that the distance between the GPU path and the
fp32 oracle is operand rounding (the GPU results are several times closer to this mode than to fp32) and to check gradients
against a tolerance that is not dominated by that rounding.  Casts are differentiable, so autograd works unchanged.
"""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F


def _r(x):
    return x.to(torch.bfloat16).to(torch.float32) if isinstance(x, torch.Tensor) and x.dtype == torch.float32 else x


@contextlib.contextmanager
def bf16_operands():
    lin, conv, mm = F.linear, F.conv1d, torch.matmul

    def linear(x, w, b=None):
        return lin(_r(x), _r(w), b)

    def conv1d(x, w, b=None, *a, **k):
        return conv(_r(x), _r(w), b, *a, **k)

    def matmul(a, b, **k):
        return mm(_r(a), _r(b), **k)

    F.linear, F.conv1d, torch.matmul = linear, conv1d, matmul
    try:
        yield
    finally:
        F.linear, F.conv1d, torch.matmul = lin, conv, mm
