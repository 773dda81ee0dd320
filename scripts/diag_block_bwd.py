#!/usr/bin/env python
"""Bottleneck backward: bf16 engine (tcgen05) vs bf16 engine on the SIMT kernels vs fp32 engine, same bf16-rounded
inputs and weights.  Localises a bf16-path bug to the tensor-core kernels or to the bf16 row-streaming kernels."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adamml_b200 import ops  # noqa: E402
from adamml_b200.engine import Exec  # noqa: E402

dev = torch.device("cuda:0")
_Block = importlib.import_module("adamml_b200.models.resnet")._Block


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max()).item()


def rms(a, b):
    a, b = a.float(), b.float()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()


for cfg in [(256, 128, 2, True), (512, 128, 1, False), (64, 64, 1, True)]:
    inpl, planes, stride, ds = cfg
    g = torch.Generator().manual_seed(0)
    blk = _Block(inpl, planes, stride, True, ds)
    with torch.no_grad():
        for m in blk.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.3)
            if isinstance(m, torch.nn.Conv2d):
                m.weight.copy_(m.weight.bfloat16().float())
    blk = blk.to(dev).train()
    G, ipg, H = 2, 6, 28
    x = torch.randn(G * ipg, inpl, H, H, generator=g).to(dev).bfloat16().float()
    res = {}
    for tag, dt, mode in (("fp32", torch.float32, "auto"), ("bf16-simt", torch.bfloat16, "simt"), ("bf16-tc", torch.bfloat16, "auto")):
        ops.TC_MODE = mode
        ex = Exec(dt, True, G, save=True)
        out = ex.bottleneck(nhwc(x).to(dt), blk)
        if "dy" not in res:
            res["dy"] = torch.randn(out.shape, generator=torch.Generator().manual_seed(1)).to(dev).bfloat16().float()
        dx = ex.bottleneck_bwd(res["dy"].to(dt).clone())
        res[tag] = (out.float(), dx.float(), {k: ex.grads[p].float() for k, p in blk.named_parameters()})
    ops.TC_MODE = "auto"
    for tag in ("bf16-simt", "bf16-tc"):
        o, d, gr = res[tag]
        o0, d0, g0 = res["fp32"]
        worst = max(gr, key=lambda k: rel(gr[k], g0[k]))
        print(f"{cfg} {tag:10s} out max {rel(o, o0):.3e} rms {rms(o, o0):.3e} | dx max {rel(d, d0):.3e} rms {rms(d, d0):.3e} | "
              f"worst grad {worst} {rel(gr[worst], g0[worst]):.3e}")
