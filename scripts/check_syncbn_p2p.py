#!/usr/bin/env python
"""torchrun --nproc-per-node 2 scripts/check_syncbn_p2p.py
sync-BN statistics over NVLink peer memory (csrc/p2p.cu) vs the NCCL all-reduce path on the same inputs:
logits / gradients must agree to fp64-summation-order noise, and both must equal a single-process run on the
concatenated batch (equal per-rank counts => identical statistics)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import O, compare_grads, namespace  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
import datetime  # noqa: E402
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=60))
from adamml_b200.dist_utils import P2PStats  # noqa: E402
from adamml_b200.models import build_model  # noqa: E402

case = dict(kind="adamml", modality=["rgb", "sound"], N=2, S=2, hw=64, training=True)
cfg = O.make_cfg(case["modality"], num_segments=2)
xs_all, y_all = O.make_inputs(cfg, 2 * world, 2, hw=64)       # global batch; rank r takes clips [2r, 2r+2)
noise = O.draw_noise(1, cfg, 2 * world, 2, True)


def run(p2p, sync, lo, hi, dtype=torch.float32):
    os.environ["ADAMML_B200_SYNCBN_P2P"] = "1" if p2p else "0"
    model, _ = build_model(namespace(case, compute_dtype=dtype))
    model.load_state_dict(O.fill_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=0))
    if sync:
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    model = model.to(dev).train()
    n = hi - lo
    M = 2
    expo = torch.stack(noise["expo"]).view(2, M, 2 * world, 2)[:, :, lo:hi].reshape(2, M * n, 2).to(dev)
    drop = [torch.cat([noise["drop"][s][m].view(2 * world, -1)[lo:hi] for s in range(2)], 0).to(dev) for m in range(M)]
    logits, dec = model([x[lo:hi].to(dev) for x in xs_all], noise=dict(expo=expo, drop=drop))
    # loss normalised by the GLOBAL batch so that per-rank gradients sum to the single-process gradient
    loss = F.cross_entropy(logits, y_all[lo:hi].to(dev), reduction="sum") / (2 * world)
    loss.backward()
    torch.cuda.synchronize()
    g = {k: p.grad.clone() for k, p in model.named_parameters()}
    for v in g.values():
        dist.all_reduce(v)
    return logits.detach(), g, model


lo, hi = 2 * rank, 2 * rank + 2
l_nccl, g_nccl, _ = run(False, True, lo, hi)
l_p2p, g_p2p, _ = run(True, True, lo, hi)
ws = P2PStats.get(dist.group.WORLD, dev)
assert ws is not None, "symmetric memory unavailable"
ws.check()
os.environ["ADAMML_B200_SYNCBN_P2P"] = "0"
rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()  # noqa: E731
e1 = rel(l_p2p, l_nccl)
e2 = max(rel(g_p2p[k], g_nccl[k]) for k in g_nccl)
msg = f"rank {rank}: p2p vs nccl logits {e1:.2e}, worst grad {e2:.2e}"
if rank == 0:  # single-process reference on the concatenated batch
    model, _ = build_model(namespace(case, compute_dtype=torch.float32))
    model.load_state_dict(O.fill_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=0))
    model = model.to(dev).train()
    n = 2 * world
    expo = torch.stack(noise["expo"]).to(dev)
    drop = [torch.cat([noise["drop"][s][m] for s in range(2)], 0).to(dev) for m in range(2)]
    logits, _ = model([x.to(dev) for x in xs_all], noise=dict(expo=expo, drop=drop))
    (F.cross_entropy(logits, y_all.to(dev), reduction="sum") / n).backward()
    e3 = rel(l_p2p, logits[lo:hi].detach())
    # (robust to parameters whose true gradient is exactly zero, see tests/util.py)
    bad = compare_grads(g_p2p, {k: p.grad for k, p in model.named_parameters()}, tol=1e-2)
    msg += f"; vs single-process concatenated batch: logits {e3:.2e}, {len(bad)} of {len(g_p2p)} grads beyond 1e-2"
    assert e3 < 1e-4 and len(bad) <= len(g_p2p) // 50, (msg, bad[:5])
print(msg, flush=True)
assert e1 < 1e-5 and e2 < 1e-3, msg
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
