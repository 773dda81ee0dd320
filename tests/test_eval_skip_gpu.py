"""Inference with decision-driven skipping (reference validate loop utils/utils.py:427-507).

The reference runs every main backbone on every (segment, video) pair and multiplies the logits by the policy's
0/1 decision (adamml.py:81-86, joint_resnet_mobilenetv2.py:94).  In eval mode under no_grad the product runs the
main backbones only on the selected pairs; the result must equal the run-everything path on the same decisions.
"""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import namespace  # noqa: E402

pytestmark = pytest.mark.gpu

HW = 64


def _model(dev, modality, dtype, seed=0, **over):
    from adamml_b200.models import build_model
    case = dict(kind="adamml", modality=modality, S=3)
    torch.manual_seed(seed)
    model, _ = build_model(namespace(case, compute_dtype=dtype, **over))
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():  # non-trivial BN affine / running statistics
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) * 0.5 + 0.75)
    return model.to(dev).eval()


def _inputs(dev, modality, N, S):
    ch = dict(rgb=3, flow=10, rgbdiff=15)
    g = torch.Generator(device=dev).manual_seed(5)
    xs = []
    for m in modality:
        if m == "sound":
            xs.append(torch.randn(N, S, 256, 256, device=dev, generator=g) * 3 - 5)
        else:
            xs.append(torch.randn(N, S * 8 * ch[m], HW, HW, device=dev, generator=g))
    return xs


@pytest.mark.parametrize("mode", ["device", "host"])
@pytest.mark.parametrize("dtype", ["x2", torch.float32, torch.bfloat16])
@pytest.mark.parametrize("modality", [["rgb", "sound"], ["rgb", "sound", "flow", "rgbdiff"]])
def test_selected_only_equals_run_everything(cuda, modality, dtype, mode):
    """mode "device": compaction / gather / work limit / scatter on the device, no host sync (csrc/gating.cu);
    mode "host": decisions read back, data-dependent batch shape."""
    N, S = 5, 3
    model = _model(cuda, modality, dtype)
    model.skip_mode = mode
    M = model.num_modality
    xs = _inputs(cuda, modality, N, S)
    g = torch.Generator(device=cuda).manual_seed(11)
    expo = torch.empty(S, M * N, 2, device=cuda).exponential_(generator=g)
    with torch.no_grad():
        model.skip_unselected = False
        ref_logits, ref_dec = model(xs, noise=dict(expo=expo))
        model.skip_unselected = True
        logits, dec = model(xs, noise=dict(expo=expo))
    frac = model.last_selected_fraction
    print(f"{modality} {dtype}: selected fraction {frac:.2f}")
    assert torch.equal(dec, ref_dec)
    assert 0.0 < frac < 1.0, "degenerate decisions: the test would not exercise the gather/scatter"
    assert abs(frac - dec.mean().item()) < 1e-6
    err = ((logits - ref_logits).abs().max() / ref_logits.abs().max()).item()
    print(f"max |skip - full| / max |full| = {err:.3e}; identical = {torch.equal(logits, ref_logits)}")
    assert err < 1e-5, err


@pytest.mark.parametrize("mode", ["device", "host"])
@pytest.mark.parametrize("thr,expect", [(1.5, 0.0), (-1.0, 1.0), (0.5, None)])
def test_rng_policy_extremes(cuda, thr, expect, mode):
    """rng_policy (adamml.py:38-40,76-78): nothing selected -> exactly zero logits without any main launch;
    everything selected -> the plain path; mixed -> equals run-everything on the same random decisions."""
    from adamml_b200 import _lib
    N, S = 4, 3
    modality = ["rgb", "sound"]
    model = _model(cuda, modality, torch.bfloat16, rng_policy=True, rng_threshold=thr)
    model.skip_mode = mode
    xs = _inputs(cuda, modality, N, S)
    with torch.no_grad():
        model.skip_unselected = False
        torch.manual_seed(3)
        n0 = _lib.launch_count()
        ref_logits, ref_dec = model(xs)
        full_launches = _lib.launch_count() - n0
        model.skip_unselected = True
        torch.manual_seed(3)
        n0 = _lib.launch_count()
        logits, dec = model(xs)
        launches = _lib.launch_count() - n0
    assert torch.equal(dec, ref_dec)
    if expect is not None:
        assert model.last_selected_fraction == expect
    if expect == 0.0:
        assert logits.abs().max() == 0 and ref_logits.abs().max() == 0
        if mode == "host":  # (device mode issues the same launches; their tile loops / threads find nothing live)
            assert launches < 0.1 * full_launches, (launches, full_launches)
    err = ((logits - ref_logits).abs().max() / ref_logits.abs().max().clamp_min(1e-30)).item()
    assert err < 1e-5, err


def test_training_or_grad_mode_never_skips(cuda):
    """Train-mode BN couples the clips of a batch and the straight-through gradient needs every backbone pass:
    skipping is confined to eval + no_grad."""
    model = _model(cuda, ["rgb", "sound"], torch.bfloat16)
    assert model.skip_unselected
    with torch.no_grad():
        assert model._can_skip()
    assert not model._can_skip()                      # grad mode on
    model.train()
    with torch.no_grad():
        assert not model._can_skip()                  # batch-statistics BN
    model.eval()
    model.main_net.nets[0].bn1.train()                # a single BN left in train mode is enough to disable it
    with torch.no_grad():
        assert not model._can_skip()


def test_device_skip_pass_is_graph_capturable(cuda):
    """NS3: with the gating on the device nothing in the inference pass depends on the host knowing the decisions, so
    data layer + policy + compaction + gated main nets + fusion capture into ONE CUDA graph; replays on new inputs /
    new Gumbel noise must reproduce the eager pass bit for bit (different selections, same launches)."""
    N, S = 4, 3
    modality = ["rgb", "sound"]
    model = _model(cuda, modality, "x2")
    model.skip_mode = "device"
    M = model.num_modality
    xs = _inputs(cuda, modality, N, S)
    g = torch.Generator(device=cuda).manual_seed(21)
    expo = torch.empty(S, M * N, 2, device=cuda).exponential_(generator=g)
    with torch.no_grad():
        for _ in range(2):  # warm-up (lazy allocations, bn key assignment)
            model(xs, noise=dict(expo=expo))
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            g_logits, g_dec = model(xs, noise=dict(expo=expo))
        seen = set()
        for trial in range(3):
            for x in xs:
                x.copy_(torch.randn(x.shape, device=cuda, generator=g) * (1 + trial))
            expo.copy_(torch.empty_like(expo).exponential_(generator=g))
            graph.replay()
            torch.cuda.synchronize()
            r_logits, r_dec = g_logits.clone(), g_dec.clone()
            e_logits, e_dec = model(xs, noise=dict(expo=expo))
            assert torch.equal(r_dec, e_dec)
            assert torch.equal(r_logits, e_logits), (r_logits - e_logits).abs().max()
            seen.add(tuple(r_dec.flatten().tolist()))
        assert len(seen) > 1, "the replays never changed the selection: the test would not exercise the gating"
