"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: packed sync-BN statistic exchange, max-over-ranks
timing, batch sharding.  The N>1 data path itself (DDP + NCCL) is exercised by bench.py --gpus N on the GPU box."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adamml_b200.dist_utils import allreduce_grads, allreduce_stats, max_over_ranks, shard_batch, sync_bn_group
    try:
        G, C, per_rank = 3, 5, 7
        g = torch.Generator().manual_seed(0)
        z = torch.randn(world, G, per_rank, C, generator=g, dtype=torch.float64)  # same on all ranks
        mine = z[rank]
        sums = torch.stack([mine.sum(1), (mine * mine).sum(1)], -1)                # [G, C, 2]
        bn = torch.nn.SyncBatchNorm(C)
        pg = sync_bn_group(bn)
        assert pg is not None
        assert sync_bn_group(torch.nn.BatchNorm2d(C)) is None
        count = allreduce_stats(sums, per_rank, pg)
        allz = z.permute(1, 0, 2, 3).reshape(G, world * per_rank, C)
        want = torch.stack([allz.sum(1), (allz * allz).sum(1)], -1)
        assert count == world * per_rank
        assert torch.allclose(sums, want, rtol=1e-12, atol=1e-12)
        # statistics of the concatenated batch == single-process BatchNorm on that batch
        mean = sums[..., 0] / count
        var = sums[..., 1] / count - mean * mean
        assert torch.allclose(mean, allz.mean(1)) and torch.allclose(var, allz.var(1, unbiased=False))
        assert max_over_ranks(10.0 + rank) == 10.0 + world - 1
        assert shard_batch(72, world) == 72 // world
        # flat gradient averaging == DDP semantics (mean over ranks; parameters without grad are skipped)
        ps = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(1))]
        ps[0].grad = torch.full((3, 2), float(rank + 1))
        ps[1].grad = torch.arange(5.0) * (rank + 1)
        allreduce_grads(ps)
        mean = sum(range(1, world + 1)) / world
        assert torch.allclose(ps[0].grad, torch.full((3, 2), mean)) and torch.allclose(ps[1].grad, torch.arange(5.0) * mean)
        assert ps[2].grad is None
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sync_bn_stats_and_timing_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
