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
              causality_modeling=case.get("causality", "lstm"), rng_policy=False, rng_threshold=0.5, unimodality_pretrained=[],
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


def compare_grads(named_product, named_ref, tol=2e-2, floor_frac=0.02):
    """Element-wise gradient comparison that is robust to mathematically-zero gradients.

    Some parameters have an exactly-zero true gradient (e.g. the beta of MobileNetV2's linear
    bottleneck BN: it only shifts the input of a 1x1 conv that is followed by a train-mode BN, which
    removes any per-channel constant).  Both sides then hold round-off noise, so the error of every
    tensor is measured against max(max|ref|, floor_frac * median max|ref| of its sub-network).
    Returns the list of (name, err, max|ref|) that exceed `tol`.
    """
    import statistics
    groups = {}
    for k, r in named_ref.items():
        parts = k.split(".")
        gk = ".".join(parts[:parts.index("nets") + 2]) if "nets" in parts else parts[0]
        groups.setdefault(gk, []).append(r.detach().abs().max().item())
    med = {gk: statistics.median(v) for gk, v in groups.items()}
    bad = []
    for k, r in named_ref.items():
        parts = k.split(".")
        gk = ".".join(parts[:parts.index("nets") + 2]) if "nets" in parts else parts[0]
        p = named_product[k]
        r64, p64 = r.detach().double().cpu(), p.detach().double().cpu()
        den = max(r64.abs().max().item(), floor_frac * med[gk], 1e-30)
        err = (p64 - r64).abs().max().item() / den
        if not err < tol:
            bad.append((k, err, r64.abs().max().item()))
    return bad


def grad_errors(named, truth, floor_frac=0.02):
    """Per-parameter normalised max error against a (float64) truth -> (mean, max, worst name)."""
    import statistics
    def gkey(k):
        parts = k.split(".")
        return ".".join(parts[:parts.index("nets") + 2]) if "nets" in parts else parts[0]
    groups = {}
    for k, r in truth.items():
        groups.setdefault(gkey(k), []).append(r.detach().abs().max().item())
    med = {g: statistics.median(v) for g, v in groups.items()}
    errs = {}
    for k, r in truth.items():
        r64, p64 = r.detach().double().cpu(), named[k].detach().double().cpu()
        den = max(r64.abs().max().item(), floor_frac * med[gkey(k)], 1e-30)
        errs[k] = (p64 - r64).abs().max().item() / den
    worst = max(errs, key=errs.get)
    return sum(errs.values()) / len(errs), errs[worst], worst


def assert_grads_as_good_as_reference(prod, ref32, truth64, mean_factor=2.0, max_factor=4.0, mean_floor=1e-4,
                                      max_floor=1e-3):
    """Training gradients of tiny-batch BatchNorm nets are ill-conditioned: the reference's own fp32
    gradients differ from the float64 truth by percents (scripts/diag_grad_sensitivity.py: CPU fp32 mean
    2.5 % / worst 11 %, torch-CUDA fp32 3.5 % / 25 % on resnet50_rgb_b2).  The bar is therefore: the
    product's error against the float64 oracle must be of the same size as the reference's fp32 error."""
    pm, px, pk = grad_errors(prod, truth64)
    rm, rx, rk = grad_errors(ref32, truth64)
    print(f"grad error vs float64 truth: product mean {pm:.3e} max {px:.3e} ({pk}); "
          f"reference-fp32 mean {rm:.3e} max {rx:.3e} ({rk})")
    assert pm <= mean_factor * rm + mean_floor, (pm, rm)
    assert px <= max_factor * rx + max_floor, (px, rx, pk)


def oracle_run(case, seed, dtype, sd0):
    """Runs the oracle (fp32 = bit-identical to the reference, or float64 = truth) -> logits, dec, grads, sd."""
    import torch.nn.functional as F
    cfg = O.make_cfg(case["modality"], num_segments=case["S"], causality_modeling=case.get("causality", "lstm"))
    N, S_run, training = case["N"], case.get("S_run", case["S"]), case["training"]
    xs, y = O.make_inputs(cfg, N, S_run, hw=case["hw"])
    sd = {}
    for k, v in sd0.items():
        t = v.detach().clone()
        if t.is_floating_point():
            t = t.to(dtype)
            if not k.endswith(("running_mean", "running_var")):
                t.requires_grad_(True)
        sd[k] = t
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        if case["kind"] == "resnet":
            gen = torch.Generator(); gen.manual_seed(seed)
            mask = torch.empty(N, 2048, dtype=torch.float32).bernoulli_(0.5, generator=gen).div_(0.5).to(dtype)
            logits = O.resnet_forward(sd, "", xs[0].to(dtype), cfg, training, mask)
            dec = None
            loss = F.cross_entropy(logits, y)
        else:
            noise = O.draw_noise(seed, cfg, N, S_run, training)
            noise = dict(expo=[e.to(dtype) for e in noise["expo"]],
                         drop=[[m.to(dtype) for m in per] for per in noise["drop"]])
            logits, dec = O.adamml_forward(sd, [x.to(dtype) for x in xs], cfg, training, noise, num_segments=S_run)
            loss = F.cross_entropy(logits, y)
            if training:
                loss = loss + O.policy_loss(dec, [1.0] * dec.shape[-1], 10.0, logits, y)
        if training:
            loss.backward()
    finally:
        torch.set_default_dtype(old)
    grads = {k: v.grad for k, v in sd.items() if v.is_floating_point() and v.grad is not None}
    return logits.detach(), dec.detach() if dec is not None else None, grads, sd
