"""Full-size (BASELINE.json configs[1]: RGB+Audio, N=72 clips, S=5, 8 frames, 224^2) checks through
size-independent properties — the CPU oracle cannot run this size in seconds.

* eval mode: every clip is independent (running-stat BatchNorm, per-clip LSTM), so the batched launch over all
  72 clips x 5 segments must reproduce, clip for clip, what two half-batches produce (tiles, TMA boxes and BN
  groups are cut differently, the arithmetic per output element is not);
* train mode: one forward+backward at full size: finite loss/gradients for every parameter, per-segment BN
  bookkeeping (num_batches_tracked += S), decisions in {0,1}.
"""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import namespace  # noqa: E402

pytestmark = pytest.mark.gpu

CASE = dict(kind="adamml", modality=["rgb", "sound"], N=72, S=5, hw=224, training=False)


def _inputs(dev, N, S):
    g = torch.Generator(device=dev).manual_seed(123)
    rgb = torch.randn(N, S * 8 * 3, 224, 224, device=dev, generator=g)
    snd = torch.randn(N, S, 256, 256, device=dev, generator=g) * 3 - 5
    expo = torch.empty(S, 2, N, 2, device=dev).exponential_(generator=g)  # [S, M, N, 2]
    y = torch.randint(0, 31, (N,), device=dev, generator=g)
    return rgb, snd, expo, y


def _model(dev, seed=0):
    from adamml_b200.models import build_model
    torch.manual_seed(seed)
    model, _ = build_model(namespace(CASE, compute_dtype=torch.bfloat16))
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():  # non-trivial BN affine / running statistics so that folding bugs are visible
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) * 0.5 + 0.75)
    return model.to(dev)


def test_eval_full_batch_equals_half_batches(cuda):
    N, S = CASE["N"], CASE["S"]
    model = _model(cuda).eval()
    rgb, snd, expo, _ = _inputs(cuda, N, S)
    with torch.no_grad():
        full_logits, full_dec = model([rgb, snd], noise=dict(expo=expo.reshape(S, 2 * N, 2)))
        parts = []
        for lo, hi in ((0, N // 2), (N // 2, N)):
            e = expo[:, :, lo:hi].reshape(S, 2 * (hi - lo), 2).contiguous()
            parts.append(model([rgb[lo:hi].contiguous(), snd[lo:hi].contiguous()], noise=dict(expo=e)))
    logits = torch.cat([p[0] for p in parts])
    dec = torch.cat([p[1] for p in parts])
    assert full_logits.shape == (N, 31) and full_dec.shape == (N, S, 2)
    assert torch.isfinite(full_logits).all()
    assert set(full_dec.unique().tolist()) <= {0.0, 1.0}
    assert torch.equal(dec, full_dec), "policy selections changed with the batch split"
    err = (logits - full_logits).abs().max() / full_logits.abs().max()
    assert err < 1e-5, f"per-clip logits changed with the batch split: {err:.3e}"


def test_train_step_full_size(cuda):
    N, S = CASE["N"], CASE["S"]
    model = _model(cuda).train()
    rgb, snd, expo, y = _inputs(cuda, N, S)
    logits, dec = model([rgb, snd])
    loss = F.cross_entropy(logits, y) + (dec.mean(1) ** 2).mean()
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss)
    assert set(dec.detach().unique().tolist()) <= {0.0, 1.0}
    missing = [k for k, p in model.named_parameters() if p.grad is None]
    assert not missing, missing[:5]
    bad = [k for k, p in model.named_parameters() if not torch.isfinite(p.grad).all()]
    assert not bad, bad[:5]
    nz = sum(int(p.grad.abs().max() > 0) for p in model.parameters())
    assert nz > 0.95 * sum(1 for _ in model.parameters())
    for k, b in model.named_buffers():
        if k.endswith("num_batches_tracked"):
            assert int(b) == S, (k, int(b))  # one BatchNorm update per segment call (adamml.py:84-86)
