#!/usr/bin/env python
"""ONE launch of the training-stem pooling kernels at the N=72 shape, for
  ncu --set full --clock-control none --import-source on -k regex:"pool" -o gpurun_out/r2_prof_pool python scripts/ncu_pool.py"""
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
y, pos = ops.bn_act_maxpool_fwd(z, ss, G, ops.ACT_RELU)
dy = torch.randn(2880, 56, 56, 64, device=dev).bfloat16()
dx = ops.maxpool_bwd(None, dy, pos=pos, x_shape=(2880, 112, 112, 64))
torch.cuda.synchronize()
print("done")
