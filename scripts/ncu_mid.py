#!/usr/bin/env python
"""ONE launch of each x2 GEMM / conv / weight-gradient shape that sits furthest from its bound in the N=72 step
(profiles/r2_calls_N72_x2.jsonl), for
  ncu --set full --clock-control none --import-source on -k regex:"tc_" -o gpurun_out/r2_prof_mid python scripts/ncu_mid.py
digest with scripts/ncu_summary.py / scripts/ncu_hotspots.py."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adamml_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
G = 5


def rnd(*shape):
    return torch.randn(*shape, device=dev).to(BF)


def x2rand(*shape):
    t = torch.randn(*shape, device=dev)
    hi = t.bfloat16()
    return ops.X2(hi, (t - hi.float()).half())


which = sys.argv[1] if len(sys.argv) > 1 else "all"
for M, N, K in [(141120, 1024, 256), (1128960, 512, 128), (141120, 256, 1024), (17640, 2048, 512)]:
    if which not in ("all", "gemm"):
        break
    a = x2rand(M, 1, 1, K)
    w = ops.pack_weight(torch.randn(N, K, 1, 1, device=dev) / K ** 0.5, ops.PREC_X2)
    st = torch.empty(G, N, 2, device=dev, dtype=torch.float64)
    ops.conv_fwd(a, w, 1, 0, stats=st, rows_per_group=M // G)
    del a
if which in ("all", "conv"):
    x3 = x2rand(2880, 56, 56, 64)
    w3 = ops.pack_weight(torch.randn(64, 64, 3, 3, device=dev) / 24, ops.PREC_X2)
    st3 = torch.empty(G, 64, 2, device=dev, dtype=torch.float64)
    ops.conv_fwd(x3, w3, 1, 1, stats=st3, rows_per_group=2880 * 56 * 56 // G)
    del x3
if which in ("all", "wgrad"):
    x, dy = rnd(2880, 56, 56, 64), rnd(2880, 56, 56, 64)
    dw = torch.empty(64, 3, 3, 64, device=dev, dtype=torch.float32)
    _lib.call("tc_wgrad_bf16", x, dy, dw, 2880, 56, 56, 64, 64, 3, 3, 1, 1, 56, 56)
torch.cuda.synchronize()
print("done")
