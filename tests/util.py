"""Shared helpers for the parity tests (oracle side is test infrastructure only)."""
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle import adamml_oracle as O  # noqa: E402


def namespace(case, **over):
    """The opts.py namespace (after train_adamml.py:70-95 mutations) for a golden case."""
    mod = case["modality"]
    ns = dict(groups=8, frames_per_group=4, num_segments=case["S"], depth=50, num_classes=31, dropout=0.5,
              pooling_method="max", without_t_stride=False, fusion_point="logits", learnable_lf_weights=True,
              causality_modeling="lstm", rng_policy=False, rng_threshold=0.5, unimodality_pretrained=[],
              imagenet_pretrained=False, dataset="kinetics-sounds", dense_sampling=False, lr_scheduler="cosine",
              sync_bn=False, batch_size=72, prefix="", epochs=1)
    if case["kind"] == "resnet":
        ns.update(backbone_net="resnet", modality="rgb", input_channels=3)
    else:
        ns.update(backbone_net="adamml", modality=mod, input_channels=[O.INPUT_CHANNELS[m] for m in mod])
    ns.update(over)
    return SimpleNamespace(**ns)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def fingerprint(t):
    t = t.detach().double().flatten().cpu()
    ramp = torch.linspace(-1.0, 1.0, t.numel(), dtype=torch.float64)
    return torch.stack([t.sum(), t.abs().sum(), (t * ramp).sum()])


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def noise_for_model(noise, device):
    """oracle.draw_noise() -> the `noise=` argument of adamml_b200 AdaMML.forward."""
    out = dict(expo=torch.stack(noise["expo"]).to(device))
    if noise.get("drop"):
        S, Mm = len(noise["drop"]), len(noise["drop"][0])
        out["drop"] = [torch.cat([noise["drop"][s][m] for s in range(S)], 0).to(device) for m in range(Mm)]
    return out
