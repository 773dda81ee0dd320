#!/usr/bin/env python
"""Launches of narrow / shallow MobileNetV2 pointwise GEMM shapes, for an `ncu --set full` capture."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adamml_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
for (M, N, K, stats) in ((5898240, 32, 16, False), (9216000, 96, 16, True)):
    A = torch.randn(M, K, device=dev).to(BF)
    B = torch.randn(N, K, device=dev).to(BF)
    D = torch.empty(M, N, device=dev, dtype=BF)
    st = torch.empty(5, N, 2, device=dev, dtype=torch.float64) if stats else None
    for _ in range(2):
        _lib.call("tc_gemm_bf16", A, B, D, M, N, K, 0, 0, 0, _lib.BF16, st, M // 5 if stats else 0)
    torch.cuda.synchronize()
print("done")
