"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

For every case: build the reference model (HTTP weight download stubbed), load the
deterministic parameters of oracle.fill_state_dict, run forward(+backward) under
torch.manual_seed(seed) and store logits / decisions / loss / gradient fingerprints /
post-step BN running statistics.  It also asserts that the oracle restatement reproduces the
reference on the same seed (this is the pin of the oracle).
"""
import argparse
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import adamml_oracle as O  # noqa: E402

CASES = {
    # name: (kind, modality, N, S(model), S(run), hw, training)
    "resnet50_rgb_b2": dict(kind="resnet", modality=["rgb"], N=2, S=1, hw=224, training=True),
    "adamml_rgb_sound_train": dict(kind="adamml", modality=["rgb", "sound"], N=2, S=2, hw=224, training=True),
    "adamml_rgb_sound_eval": dict(kind="adamml", modality=["rgb", "sound"], N=2, S=2, S_run=3, hw=96, training=False),
    "adamml_rgb_flow_train": dict(kind="adamml", modality=["rgb", "flow", "rgbdiff"], N=1, S=2, hw=96, training=True),
    "adamml_rgb_sound_flow_train": dict(kind="adamml", modality=["rgb", "sound", "flow", "rgbdiff"], N=2, S=2, hw=64,
                                        training=True),
    # causality_modeling=None: per-segment FC policy, one Gumbel draw over all rows (policy_net.py:330-339)
    "adamml_rgb_sound_nocausal_train": dict(kind="adamml", modality=["rgb", "sound"], N=2, S=2, hw=64, training=True,
                                            causality=None),
    # the benchmark's own segment counts: S = 5 in training (5-step LSTM recurrence, 5 sequential running-stat updates,
    # README.md:89-95) and num_segments = val_num_clips = 10 in validation (utils/utils.py:458, opts.py:122)
    "adamml_rgb_sound_train_s5": dict(kind="adamml", modality=["rgb", "sound"], N=2, S=5, hw=96, training=True),
    "adamml_rgb_sound_eval_s10": dict(kind="adamml", modality=["rgb", "sound"], N=2, S=5, S_run=10, hw=96,
                                      training=False),
}


def import_reference():
    sys.path.insert(0, "/root/reference")
    import models  # noqa
    from models import policy_net
    policy_net.MobileNetV2.load_imagenet_model = lambda self: None  # policy_net.py:193-203 (no network)
    return models


def build_reference(models, case):
    from types import SimpleNamespace
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
    model, arch = models.build_model(SimpleNamespace(**ns))
    return model, arch


def fingerprint(t):
    """Small, order-sensitive fingerprint of a tensor: [sum, abs-sum, dot with a fixed ramp]."""
    t = t.detach().double().flatten()
    ramp = torch.linspace(-1.0, 1.0, t.numel(), dtype=torch.float64)
    return torch.stack([t.sum(), t.abs().sum(), (t * ramp).sum()])


def run_case(name, case, models):
    torch.set_num_threads(os.cpu_count())
    seed = 1
    cfg = O.make_cfg(case["modality"], num_segments=case["S"], causality_modeling=case.get("causality", "lstm"))
    model, arch = build_reference(models, case)
    ref_sd = model.state_dict()
    sd = O.fill_state_dict({k: v.shape for k, v in ref_sd.items()}, seed=0)
    model.load_state_dict(sd, strict=True)
    training = case["training"]
    model.train(training)
    N, S_run = case["N"], case.get("S_run", case["S"])
    xs, y = O.make_inputs(cfg, N, S_run, hw=case["hw"])
    out = {"case": dict(case), "arch": arch, "keys": sorted(ref_sd.keys()), "seed": seed}

    t0 = time.time()
    torch.manual_seed(seed)
    if case["kind"] == "resnet":
        logits = model(xs[0])
        loss = torch.nn.functional.cross_entropy(logits, y)
        decisions = None
    else:
        if training:
            logits, decisions = model(xs)
        else:
            with torch.no_grad():
                logits, decisions = model(xs, num_segments=S_run)
        loss = torch.nn.functional.cross_entropy(logits, y)
        if training:
            sys.path.insert(0, "/root/reference")
            from utils.utils import compute_policy_loss
            loss = loss + compute_policy_loss("blockdrop", decisions, [1.0] * decisions.shape[-1], 10.0, logits, y)
    grads = {}
    if training:
        loss.backward()
        grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    print(f"[{name}] reference fwd{'+bwd' if training else ''}: {time.time() - t0:.1f}s  loss={loss.item():.6f}")

    # ---- pin the oracle against the reference on the same seed ----
    osd = O.clone_sd(sd)
    if case["kind"] == "resnet":
        g = torch.Generator(); g.manual_seed(seed)
        mask = torch.empty(N, 2048).bernoulli_(0.5, generator=g).div_(0.5)
        o_logits = O.resnet_forward(osd, "", xs[0], cfg, training, mask)
        o_loss = torch.nn.functional.cross_entropy(o_logits, y)
        o_dec = None
    else:
        noise = O.draw_noise(seed, cfg, N, S_run, training)
        if training:
            o_logits, o_dec = O.adamml_forward(osd, xs, cfg, True, noise)
        else:
            with torch.no_grad():
                o_logits, o_dec = O.adamml_forward(osd, xs, cfg, False, noise, num_segments=S_run)
        o_loss = torch.nn.functional.cross_entropy(o_logits, y)
        if training:
            o_loss = o_loss + O.policy_loss(o_dec, [1.0] * o_dec.shape[-1], 10.0, o_logits, y)
    err = (o_logits - logits).abs().max().item() / logits.abs().max().item()
    print(f"[{name}] oracle vs reference: logits rel err {err:.2e}, loss {o_loss.item():.6f}")
    assert err < 1e-5, "oracle restatement does not reproduce the reference"
    if decisions is not None:
        assert torch.equal(o_dec, decisions.detach()), "oracle decisions differ from the reference"
    if training:
        o_loss.backward()
        worst = 0.0
        for k, gref in grads.items():
            go = osd[k].grad
            assert go is not None, k
            e = (go - gref).abs().max().item() / max(gref.abs().max().item(), 1e-12)
            worst = max(worst, e)
        print(f"[{name}] oracle vs reference: worst grad rel err {worst:.2e} over {len(grads)} tensors")
        assert worst < 2e-3
        new_sd = model.state_dict()
        for k in new_sd:
            if k.endswith(("running_mean", "running_var")):
                e = (osd[k] - new_sd[k]).abs().max().item() / max(new_sd[k].abs().max().item(), 1e-12)
                assert e < 1e-5, (k, e)

    out["logits"] = logits.detach().clone()
    out["loss"] = loss.detach().clone()
    out["decisions"] = decisions.detach().clone() if decisions is not None else None
    out["grad_fp"] = {k: fingerprint(v) for k, v in grads.items()}
    small = [k for k, v in grads.items() if v.numel() <= 4096]
    out["grad_small"] = {k: grads[k].detach().clone() for k in small}
    new_sd = model.state_dict()
    out["running_fp"] = {k: fingerprint(v) for k, v in new_sd.items() if k.endswith(("running_mean", "running_var"))}
    out["num_batches_tracked"] = {k: int(v) for k, v in new_sd.items() if k.endswith("num_batches_tracked")}
    torch.save(out, os.path.join(HERE, name + ".pt"))
    print(f"[{name}] saved ({os.path.getsize(os.path.join(HERE, name + '.pt')) / 1e3:.0f} KB)")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*", default=list(CASES))
    a = ap.parse_args()
    models = import_reference()
    for n in a.cases:
        run_case(n, CASES[n], models)
