#!/usr/bin/env python
"""The "real bar" (SURVEY.md §2.2): the reference's own algorithm executed by this image's
torch / cuDNN / cuBLAS on the same B200 — the oracle restatement (bit-identical to the reference on CPU,
tests/test_oracle_golden.py) run on CUDA tensors, so every conv / BN / pool / LSTM op dispatches to the library
kernels the reference's nn.Modules would use, inside the reference's Python segment loops.

Modes: "tf32" = torch defaults (cudnn.allow_tf32 = True: what `python train_adamml.py` runs on an Ampere+ GPU),
"fp32" = TF32 off (true fp32 FFMA convs), "bf16_cl" = autocast(bfloat16) + channels_last weights/inputs (the
stronger courtesy baseline).  Reports clips/s (fwd + CE/policy loss + bwd + Adam(policy) + SGD(main)) at the largest
batch that fits, and each mode's logits error / selection mismatches against the committed CPU goldens.

    python scripts/bench_gpu_reference.py [--batch 72] [--steps 5] > profiles/r2_gpu_reference.log
"""
import argparse
import contextlib
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import O, load_golden, namespace, rel  # noqa: E402

dev = torch.device(os.environ.get("ADAMML_REF_DEVICE", "cuda:0"))


def set_mode(mode):
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    torch.backends.cuda.matmul.allow_tf32 = False  # torch default
    torch.backends.cudnn.benchmark = True
    if mode == "bf16_cl":
        return lambda: torch.autocast("cuda", dtype=torch.bfloat16)
    return contextlib.nullcontext


def shapes_for(case):
    from adamml_b200.models import build_model
    model, _ = build_model(namespace(case))
    return {k: v.shape for k, v in model.state_dict().items()}


def to_dev(sd0, mode, grad=True):
    sd = {}
    for k, v in sd0.items():
        t = v.detach().clone().to(dev)
        if mode == "bf16_cl" and t.dim() == 4:
            t = t.contiguous(memory_format=torch.channels_last)
        if grad and t.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            t.requires_grad_(True)
        sd[k] = t
    return sd


def golden_errors(mode):
    ctx = set_mode(mode)
    out = {}
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import CASES
    for name, case in CASES.items():
        if case["kind"] != "adamml":
            continue
        g = load_golden(name)
        cfg = O.make_cfg(case["modality"], num_segments=case["S"], causality_modeling=case.get("causality", "lstm"))
        N, S_run, training = case["N"], case.get("S_run", case["S"]), case["training"]
        xs, _ = O.make_inputs(cfg, N, S_run, hw=case["hw"])
        noise = O.draw_noise(1, cfg, N, S_run, training)
        noise = dict(expo=[e.to(dev) for e in noise["expo"]], drop=[[m.to(dev) for m in per] for per in noise["drop"]])
        sd = to_dev(O.fill_state_dict(shapes_for(case), seed=0), mode, grad=False)
        with torch.no_grad(), torch.device(dev), ctx():
            logits, dec = O.adamml_forward(sd, [x.to(dev) for x in xs], cfg, training, noise, num_segments=S_run)
        flips = int((dec.float().cpu() != g["decisions"]).sum())
        out[name] = dict(logits_rel=rel(logits.float(), g["logits"]), flipped=flips, of=dec.numel())
    return out


def bench(mode, N, S, steps, warmup, modality):
    ctx = set_mode(mode)
    case = dict(kind="adamml", modality=modality, S=S)
    cfg = O.make_cfg(modality, num_segments=S)
    sd = to_dev(O.fill_state_dict(shapes_for(case), seed=0), mode)
    g = torch.Generator(device=dev).manual_seed(123)
    ch = {"rgb": 3, "flow": 10, "rgbdiff": 15}
    xs = []
    for m in modality:
        shape = (N, S, 256, 256) if m == "sound" else (N, S * 8 * ch[m], 224, 224)
        xs.append(torch.randn(shape, device=dev, generator=g))
    y = torch.randint(0, 31, (N,), device=dev, generator=g)
    p_params = [v for k, v in sd.items() if k.startswith("policy_net.") and v.requires_grad]
    m_params = [v for k, v in sd.items() if k.startswith("main_net.") and v.requires_grad]
    p_opt = torch.optim.Adam(p_params, 0.01, weight_decay=1e-4)
    opt = torch.optim.SGD(m_params, 0.01, momentum=0.9, weight_decay=1e-4)

    def step(it):
        noise = O.draw_noise(it, cfg, N, S, True)
        noise = dict(expo=[e.to(dev) for e in noise["expo"]], drop=[[m.to(dev) for m in per] for per in noise["drop"]])
        p_opt.zero_grad(set_to_none=True)
        opt.zero_grad(set_to_none=True)
        with torch.device(dev), ctx():
            inp = xs
            if mode == "bf16_cl":
                inp = list(xs)
            logits, dec = O.adamml_forward(sd, inp, cfg, True, noise)
            loss = F.cross_entropy(logits.float(), y) + O.policy_loss(dec.float(), [1.0] * dec.shape[-1], 10.0,
                                                                      logits.float(), y)
        loss.backward()
        p_opt.step()
        opt.step()
        return loss

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        step(warmup + i).item()  # the reference's loop reads the loss every iteration (utils/utils.py:384)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    wall = (time.perf_counter() - t0) / steps * 1e3
    return dict(mode=mode, batch=N, ms_per_step=ms, wall_ms_per_step=wall, clips_per_s=N / (ms / 1e3),
                peak_mem_gib=torch.cuda.max_memory_allocated() / 2 ** 30)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=72)
    ap.add_argument("--segments", type=int, default=5)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--modality", default="rgb,sound")
    ap.add_argument("--modes", default="tf32,fp32,bf16_cl")
    ap.add_argument("--no-golden", action="store_true")
    a = ap.parse_args()
    modality = a.modality.split(",")
    print(json.dumps(dict(torch=torch.__version__, cudnn=torch.backends.cudnn.version(),
                          gpu=torch.cuda.get_device_name(0))), flush=True)
    for mode in a.modes.split(","):
        if not a.no_golden:
            print(json.dumps(dict(mode=mode, golden=golden_errors(mode))), flush=True)
        N = a.batch
        while N >= 4:
            try:
                torch.cuda.empty_cache()
                torch.cuda.reset_peak_memory_stats()
                r = bench(mode, N, a.segments, a.steps, a.warmup, modality)
                print(json.dumps(r), flush=True)
                break
            except torch.OutOfMemoryError:
                print(json.dumps(dict(mode=mode, batch=N, oom=True)), flush=True)
                import gc
                gc.collect()
                N = N * 2 // 3
    print("done", flush=True)


if __name__ == "__main__":
    main()
