#!/usr/bin/env python
"""Inference throughput with decision-driven skipping: eval-mode AdaMML forward at
N=72 clips, S=10 segments (the reference's test-time setting, utils/utils.py:427-507), RGB+Audio.
Times the run-everything path (what the reference computes) and the selected-only path on the same inputs, in the
default x2 precision mode (or bf16 / with the fused inference epilogue off for comparison).
Usage: python scripts/bench_eval_skip.py [N] [S] [x2|bf16] [fuse=1|0]"""
import json
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adamml_b200.models import build_model  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 72
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    prec = sys.argv[3] if len(sys.argv) > 3 else "x2"
    from adamml_b200 import engine, ops
    if len(sys.argv) > 4:
        engine.FUSE_EVAL = sys.argv[4] != "0"
    dtype = {"x2": ops.PREC_X2, "bf16": torch.bfloat16}[prec]
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    results = []
    for rng_policy, thr in ((False, 0.5), (True, 0.25), (True, 0.5), (True, 0.75)):
        ns = SimpleNamespace(backbone_net="adamml", groups=8, frames_per_group=4, num_segments=S, depth=50,
                             num_classes=31, dropout=0.5, pooling_method="max", without_t_stride=False,
                             fusion_point="logits", learnable_lf_weights=True, causality_modeling="lstm",
                             rng_policy=rng_policy, rng_threshold=thr, unimodality_pretrained=[],
                             imagenet_pretrained=False, modality=["rgb", "sound"], input_channels=[3, 1],
                             dataset="kinetics-sounds", dense_sampling=False, lr_scheduler="cosine", sync_bn=False,
                             batch_size=N, prefix="", epochs=1,
                             compute_dtype=dtype)
        torch.manual_seed(0)
        model, _ = build_model(ns)
        model = model.to(dev).eval()
        g = torch.Generator(device=dev).manual_seed(123)
        rgb = torch.randn(N, S * 24, 224, 224, device=dev, generator=g)
        snd = torch.randn(N, S, 256, 256, device=dev, generator=g)

        def run(skip):
            model.skip_unselected = skip
            torch.manual_seed(1)
            with torch.no_grad():
                return model([rgb, snd], num_segments=S)

        row = {"policy": "rng>%.2f" % thr if rng_policy else "lstm (random init)", "N": N, "S": S, "precision": prec,
               "fused_epilogue": engine.FUSE_EVAL}
        for skip in (False, True):
            for _ in range(3):
                run(skip)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                out = run(skip)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            row["skip" if skip else "full"] = {"ms": round(ms, 2), "clips_per_s": round(N / ms * 1e3, 1)}
            if skip:
                row["selected_fraction"] = round(model.last_selected_fraction, 3)
                row["decision_mean"] = round(out[1].mean().item(), 3)
        row["speedup"] = round(row["full"]["ms"] / row["skip"]["ms"], 3)
        row["skip_mode"] = model.skip_mode
        if model.skip_mode == "device" and not rng_policy:
            # the device-gated pass has no host sync: capture it once, replay it with one launch
            model.skip_unselected = True
            expo = model.policy_net.draw_gumbel_noise(S, N, dev)
            with torch.no_grad():
                model([rgb, snd], num_segments=S, noise=dict(expo=expo))
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    model([rgb, snd], num_segments=S, noise=dict(expo=expo))
            for _ in range(2):
                graph.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            row["skip_graph"] = {"ms": round(ms, 2), "clips_per_s": round(N / ms * 1e3, 1)}
            del graph
        print(json.dumps(row), flush=True)
        results.append(row)
        del model, rgb, snd
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
