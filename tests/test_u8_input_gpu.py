"""Decoded uint8 frames straight into the data layer.

The reference's loader turns the stacked uint8 frames into the network input with ToTorchFormatTensor
(`img.float().div(255)`, utils/video_transforms.py:321-343) and GroupNormalize (`t.sub_(m).div_(s)` per channel
plane with the mean/std repeated over the frames, :62-84).  The product accepts the uint8 tensor itself and applies
the same fp32 arithmetic inside its re-layout kernels: results must be bit-identical to normalising first.
"""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import namespace  # noqa: E402

MEAN = {"rgb": [0.485, 0.456, 0.406], "flow": [0.5]}
STD = {"rgb": [0.229, 0.224, 0.225], "flow": [sum([0.229, 0.224, 0.225]) / 3]}


def loader_normalise(x_u8, mean, std):
    """What the reference's transform pipeline does to one clip tensor [(F*C), H, W] (on the CPU, like the loader)."""
    out = []
    for clip in x_u8.cpu():
        t = clip.float().div(255)
        rep_mean = mean * (t.size(0) // len(mean))
        rep_std = std * (t.size(0) // len(std))
        for plane, m, s in zip(t, rep_mean, rep_std):
            plane.sub_(m).div_(s)
        out.append(t)
    return torch.stack(out)


@pytest.mark.parametrize("m", ["rgb", "flow"])
def test_loader_restatement_matches_reference_transforms(m):
    """not gpu: `loader_normalise` above is pinned, bit for bit, to the reference's own Stack -> ToTorchFormatTensor
    -> GroupNormalize run (tests/golden/make_golden_u8.py -> u8_normalize.pt)."""
    from util import load_golden
    gold = load_golden("u8_normalize")[m]
    assert gold["mean"] == MEAN[m] and gold["std"] == pytest.approx(STD[m], abs=0, rel=1e-15)
    got = loader_normalise(gold["u8"][None], gold["mean"], gold["std"])[0]
    assert torch.equal(got, gold["normalized"])


@pytest.mark.gpu
@pytest.mark.parametrize("m,F", [("rgb", 4), ("flow", 1)])
def test_u8_kernel_matches_reference_transform_golden(cuda, m, F):
    """The device-side normalisation reproduces the reference loader's output on its own golden vector."""
    from adamml_b200 import ops
    from util import load_golden
    gold = load_golden("u8_normalize")[m]
    x8 = gold["u8"][None].to(cuda)                                  # [1, F*C, 6, 5]
    C = x8.shape[1] // F
    norm = (torch.tensor(gold["mean"] * (C // len(gold["mean"])), device=cuda),
            torch.tensor(gold["std"] * (C // len(gold["std"])), dtype=torch.float32, device=cuda))
    y = ops.pack_frames(x8, 1, F, C, torch.float32, norm=norm)      # NHWC [F, 6, 5, C]
    ref = gold["normalized"].view(F, C, 6, 5).permute(0, 2, 3, 1).to(cuda)
    assert torch.equal(y, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("m,C", [("rgb", 3), ("flow", 10)])
def test_u8_relayout_kernels_bit_identical(cuda, m, C):
    from adamml_b200 import ops
    N, S, F, H, W = 2, 2, 8, 32, 48
    g = torch.Generator().manual_seed(3)
    x8 = torch.randint(0, 256, (N, S * F * C, H, W), generator=g, dtype=torch.uint8)
    xf = loader_normalise(x8, MEAN[m], STD[m]).to(cuda)
    x8 = x8.to(cuda)
    norm = (torch.tensor(MEAN[m] * (C // len(MEAN[m])), device=cuda), torch.tensor(STD[m] * (C // len(STD[m])), device=cuda))
    for dt in (torch.float32, torch.bfloat16):
        assert torch.equal(ops.pack_frames(x8, S, F, C, dt, norm=norm), ops.pack_frames(xf, S, F, C, dt))
        assert torch.equal(ops.resize_frames(x8, S, F, C, 20, 28, 2, dt, norm=norm),
                           ops.resize_frames(xf, S, F, C, 20, 28, 2, dt))
    a, b = ops.pack_frames_s2d(x8, S, F, C, norm=norm), ops.pack_frames_s2d(xf, S, F, C)
    assert torch.equal(a.t, b.t)
    with pytest.raises(ValueError):
        ops.pack_frames(x8, S, F, C, torch.bfloat16)          # uint8 without its normalisation


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_model_accepts_u8_frames(cuda, dtype):
    from adamml_b200.models import build_model
    case = dict(kind="adamml", modality=["rgb", "sound"], S=2)
    torch.manual_seed(0)
    model, _ = build_model(namespace(case, compute_dtype=dtype))
    model = model.to(cuda).train()
    N, S, HW = 2, 2, 64
    g = torch.Generator().manual_seed(4)
    rgb8 = torch.randint(0, 256, (N, S * 8 * 3, HW, HW), generator=g, dtype=torch.uint8)
    rgbf = loader_normalise(rgb8, MEAN["rgb"], STD["rgb"]).to(cuda)
    snd = (torch.randn(N, S, 256, 256, generator=g) * 3 - 5).to(cuda)
    expo = torch.empty(S, 2 * N, 2).exponential_(generator=g).to(cuda)
    drop = [torch.ones(S * N, 2048, device=cuda), torch.ones(S * N, 1280, device=cuda)]
    x8 = rgb8.to(cuda)
    # inference (running-statistics BN): deterministic, so the two input forms must agree bit for bit
    model.eval()
    with torch.no_grad():
        lf, df = model([rgbf, snd], noise=dict(expo=expo))
        l8, d8 = model([x8, snd], noise=dict(expo=expo))
    assert torch.equal(df, d8) and torch.equal(lf, l8)
    # training step: same operands, but the fp64 atomics of the statistics / wgrad may sum in another order
    model.train()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    outs = []
    for x in (rgbf, x8):
        model.load_state_dict(sd)
        model.zero_grad(set_to_none=True)
        logits, dec = model([x, snd], noise=dict(expo=expo, drop=drop))
        logits.square().mean().backward()
        # first convolutions consuming the frames: main ResNet stem (gated: zero if rgb was never selected) and
        # the policy MobileNetV2's first conv (always reached through the straight-through estimator)
        grads = [model.main_net.nets[0].conv1.weight.grad, model.policy_net.joint_net.nets[0].features[0][0].weight.grad]
        outs.append((logits.detach().clone(), dec.detach().clone(), [g.detach().clone() for g in grads]))
    assert torch.equal(outs[0][1], outs[1][1])
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert ((outs[0][0] - outs[1][0]).abs().max() / outs[0][0].abs().max()) < tol
    assert outs[0][2][1].abs().max() > 0
    for ga, gb in zip(outs[0][2], outs[1][2]):
        assert ((ga - gb).abs().max() / ga.abs().max().clamp_min(1e-30)) < 10 * tol
