#!/usr/bin/env python
"""ONE launch of each dominant kernel of the default-mode (x2) training step at its N=72 shape, for
  ncu --set full --clock-control none --import-source on -k regex:"tc_|bn_|dw_" -o gpurun_out/r2_prof_ops python scripts/ncu_ops_x2.py
digest with scripts/ncu_summary.py -> profiles/r2_ncu_ops_x2_N72.txt (duration, DRAM bytes = roofline.traffic, pipes)."""
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


def x2rand(*shape):
    t = torch.randn(*shape, device=dev)
    hi = t.bfloat16()
    return ops.X2(hi, (t - hi.float()).half())


G, rows = 5, 1806336
M = rows * G
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def want(k):
    return which in ("all", k)


if want("gemm") or want("bn") or want("bwd"):
    # layer1 conv3 (64 -> 256) forward, x2, with fused BN statistics
    a64 = x2rand(M, 1, 1, 64)
    w = ops.pack_weight(torch.randn(256, 64, 1, 1, device=dev) / 8, ops.PREC_X2)
    st = torch.empty(G, 256, 2, device=dev, dtype=torch.float64)
    z, _ = ops.conv_fwd(a64, w, 1, 0, stats=st, rows_per_group=rows)
if want("bn") or want("bwd"):
    # bn3 + identity + ReLU, x2
    res = x2rand(M, 1, 1, 256)
    ss = torch.rand(G, 256, 2, device=dev)
    out = ops.bn_apply(z, ss, G, 1, res=res)
    del res
if want("conv"):
    # layer1 conv2 (3x3, 64 -> 64) forward, x2
    x3 = x2rand(2880, 56, 56, 64)
    w3 = ops.pack_weight(torch.randn(64, 64, 3, 3, device=dev) / 24, ops.PREC_X2)
    st3 = torch.empty(G, 64, 2, device=dev, dtype=torch.float64)
    ops.conv_fwd(x3, w3, 1, 1, stats=st3, rows_per_group=rows)
    del x3
if want("dw"):
    # MobileNetV2 depthwise layers on TMA-staged tiles (csrc/dwconv_tma.cu): x2 stride-1 forward with fused BatchNorm
    # statistics, and the fused data + weight gradient at stride 1 and stride 2 (bf16)
    wd = ops.pack_weight_dw(torch.randn(144, 1, 3, 3, device=dev))
    xd = x2rand(1440, 40, 40, 144)
    std = torch.empty(G, 144, 2, device=dev, dtype=torch.float64)
    ops.dwconv_fwd_stats(xd, wd, 1, std, 1440 // G)
    ops.dwconv_bwd(xd.hi, rnd(1440, 40, 40, 144), wd, 1)
    del xd
    wd2 = ops.pack_weight_dw(torch.randn(96, 1, 3, 3, device=dev))
    ops.dwconv_bwd(rnd(1440, 80, 80, 96), rnd(1440, 40, 40, 96), wd2, 2)
    ops.dwconv_fwd(x2rand(1440, 80, 80, 96), wd2, 2)
if want("bwd"):
    # backward (bf16 on the hi planes): BN reduce / apply of the residual layer, weight gradient of conv3
    dout = rnd(M, 256)
    mi = torch.rand(G, 256, 2, device=dev) + 0.5
    gamma = torch.rand(256, device=dev) + 0.5
    zh, oh = z.hi.view(M, 256), out.hi.view(M, 256)
    sums = ops.bn_bwd_reduce(dout, oh, zh, mi, G, 1)
    dz, _ = ops.bn_bwd_apply(dout, oh, zh, mi, gamma, sums, G, rows, 1, True)
    dw = torch.empty(256, 1, 1, 64, device=dev, dtype=torch.float32)
    _lib.call("tc_wgrad_bf16", a64.hi.view(2880, 56, 56, 64), dz.view(2880, 56, 56, 256), dw, 2880, 56, 56, 64, 256, 1,
              1, 1, 0, 56, 56)
torch.cuda.synchronize()
print("done")
