#!/usr/bin/env python
"""Layer-by-layer drift of the bf16 engine against (a) a torch fp32 run that rounds to bf16 wherever the engine
stores bf16 ("storage emulation") and (b) pure fp32, on a randomly initialised ResNet-50 (8 clips x 8 frames,
112^2, train-mode BN).  A kernel bug shows up as a jump at one block; rounding chaos as a smooth growth."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adamml_b200 import ops  # noqa: E402
from adamml_b200.engine import Exec  # noqa: E402
from adamml_b200.models.resnet import ResNet  # noqa: E402
from adamml_b200.ops import ACT_RELU  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(3)
net = ResNet(50, 8, num_classes=31, dropout=0.5, input_channels=3, compute_dtype=torch.bfloat16).to(dev).train()
g = torch.Generator().manual_seed(3)
with torch.no_grad():
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.copy_((torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75).to(dev))
            m.bias.copy_((torch.randn(m.bias.shape, generator=g) * 0.1).to(dev))
N = 8
x = torch.randn(N, 24, 112, 112, generator=g).to(dev)


def r(t):
    return t.bfloat16().float()


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()


def rms(a, b):
    return ((a.float() - b.float()).pow(2).mean().sqrt() / b.float().pow(2).mean().sqrt()).item()


def nchw(t):
    return t.permute(0, 3, 1, 2).float()


def run_torch(rounding):
    q = r if rounding else (lambda t: t)
    outs = []
    conv = lambda m, a: q(F.conv2d(a, q(m.weight), None, m.stride, m.padding))  # noqa: E731
    bn = lambda m, z: F.batch_norm(z, None, None, m.weight, m.bias, True, 0.1, m.eps)  # noqa: E731
    with torch.no_grad():
        a = q(x.view(N * 8, 3, 112, 112))
        a = q(F.relu(bn(net.bn1, conv(net.conv1, a))))
        outs.append(("stem", a))
        a = F.max_pool2d(a, 3, 2, 1)
        frames = 8
        for li in range(4):
            for bi, blk in enumerate(getattr(net, f"layer{li + 1}")):
                idn = a
                o = q(F.relu(bn(blk.bn1, conv(blk.conv1, a))))
                o = q(F.relu(bn(blk.bn2, conv(blk.conv2, o))))
                o = bn(blk.bn3, conv(blk.conv3, o))
                if blk.downsample is not None:
                    idn = bn(blk.downsample[1], conv(blk.downsample[0], a))
                a = q(F.relu(o + idn))
                outs.append((f"layer{li + 1}.{bi}", a))
            if li < 3:
                nt, c, h, w = a.shape
                v = a.view(-1, frames, c, h, w).transpose(1, 2)
                v = F.max_pool3d(v, (3, 1, 1), (2, 1, 1), (1, 0, 0))
                a = v.transpose(1, 2).contiguous().view(-1, c, h, w)
                frames //= 2
    return outs


def run_engine():
    outs = []
    with torch.no_grad():
        ex = Exec(torch.bfloat16, True, 1, save=False)
        a = ex.cba(net.pack_input(x, 1), net.conv1, net.bn1, ACT_RELU)
        outs.append(("stem", nchw(a)))
        a = ex.maxpool(a)
        frames = 8
        for li in range(4):
            for bi, blk in enumerate(getattr(net, f"layer{li + 1}")):
                a = ex.bottleneck(a, blk)
                outs.append((f"layer{li + 1}.{bi}", nchw(a)))
            if li < 3:
                a = ex.tpool(a, frames, False)
                frames //= 2
    return outs


e, t16, t32 = run_engine(), run_torch(True), run_torch(False)
print(f"{'stage':12s} {'engine~emul max':>16s} {'rms':>10s} | {'emul~fp32 max':>14s} {'rms':>10s} | {'engine~fp32 rms':>16s}")
for (k, a), (_, b), (_, c) in zip(e, t16, t32):
    print(f"{k:12s} {rel(a, b):16.3e} {rms(a, b):10.3e} | {rel(b, c):14.3e} {rms(b, c):10.3e} | {rms(a, c):16.3e}")
