#!/usr/bin/env python
"""Times the training-stem pooling kernels at the N=72 shape (CUDA events; operands exceed L2)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adamml_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
G = 5
t = torch.randn(2880, 112, 112, 64, device=dev)
hi = t.bfloat16()
z = ops.X2(hi, (t - hi.float()).half())
del t
ss = torch.rand(G, 64, 2, device=dev)


def timeit(fn, reps=6, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timeit(lambda: ops.bn_act_maxpool_fwd(z, ss, G, ops.ACT_RELU))
by = 4.0 * z.hi.numel() + (4.0 + 1.0) * z.hi.numel() / 4
print(f"bn_act_maxpool3x3s2_fwd_x2 2880x112x112x64: {ms:.3f} ms  {by / ms / 1e6:.0f} GB/s", flush=True)
y, pos = ops.bn_act_maxpool_fwd(z, ss, G, ops.ACT_RELU)
dy = torch.randn(2880, 56, 56, 64, device=dev).bfloat16()
ms = timeit(lambda: ops.maxpool_bwd(None, dy, pos=pos, x_shape=(2880, 112, 112, 64)))
by = 2.0 * dy.numel() + pos.numel() + 2.0 * z.hi.numel()
print(f"maxpool3x3s2_bwd (positions) 2880x112x112x64: {ms:.3f} ms  {by / ms / 1e6:.0f} GB/s", flush=True)
