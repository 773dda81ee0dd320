#!/usr/bin/env python
"""Runs ONE launch of each representative kernel at its N=72 shape (after one warm-up launch) so that
`ncu --set full --launch-skip ... ` captures stay short.  Used for profiles/r1_ncu_ops_*.txt."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adamml_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16


def rnd(*shape):
    return torch.randn(*shape, device=dev).to(BF)


G, rows, C = 5, 1806336, 256
z, dout, out = rnd(rows * G, C), rnd(rows * G, C), rnd(rows * G, C)
mi = torch.rand(G, C, 2, device=dev) + 0.5
gamma = torch.rand(C, device=dev) + 0.5
ss = torch.rand(G, C, 2, device=dev)
A64, B256 = rnd(rows * G, 64), rnd(256, 64)
D = torch.empty(rows * G, 256, device=dev, dtype=BF)
st = torch.empty(G, 256, 2, device=dev, dtype=torch.float64)
x3, w3 = rnd(2880, 56, 56, 64), rnd(64, 3, 3, 64)
y3 = torch.empty(2880, 56, 56, 64, device=dev, dtype=BF)
st3 = torch.empty(G, 64, 2, device=dev, dtype=torch.float64)
dw = torch.empty(64, 1, 1, 256, device=dev, dtype=torch.float32)
for it in range(2):
    sums = ops.bn_bwd_reduce(dout, out, z, mi, G, 1)
    ops.bn_bwd_apply(dout, out, z, mi, gamma, sums, G, rows, 1, True)
    ops.bn_apply(z, ss, G, 1)
    _lib.call("tc_gemm_bf16", A64, B256, D, rows * G, 256, 64, 0, 0, 0, _lib.BF16, st, rows)
    _lib.call("tc_gemm_bf16", D, rnd(64, 256), A64, rows * G, 64, 256, 0, 0, 0, _lib.BF16, None, 0)
    _lib.call("tc_conv_bf16", x3, w3, y3, None, 2880, 56, 56, 64, 64, 3, 3, 1, 1, 56, 56, st3, 576, 0)
    _lib.call("tc_wgrad_bf16", D.view(2880, 56, 56, 256), A64.view(2880, 56, 56, 64), dw, 2880, 56, 56, 256, 64, 1, 1, 1, 0,
              56, 56)
    torch.cuda.synchronize()
print("done")
