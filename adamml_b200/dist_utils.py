"""Host-side multi-GPU helpers (one process per GPU, torch.distributed).

The hot path shards along the batch (video) axis; the only data-path collectives are the DDP gradient
all-reduce (torch's reducer, NCCL) and — with --sync-bn (train_adamml.py:125-127) — one SUM all-reduce of the
packed per-(segment, channel) BatchNorm statistics per layer and pass.  These helpers are backend agnostic so
that the logic is covered by world_size-2 gloo tests on CPU (tests/test_dist_cpu.py).
"""
import torch
import torch.distributed as dist


def sync_bn_group(bn):
    """process group of an nn.SyncBatchNorm module, or None when statistics stay local."""
    if isinstance(bn, torch.nn.SyncBatchNorm) and dist.is_available() and dist.is_initialized():
        pg = bn.process_group if bn.process_group is not None else dist.group.WORLD
        if dist.get_world_size(pg) > 1:
            return pg
    return None


def allreduce_stats(sums, count, pg):
    """In-place SUM all-reduce of packed statistics [G, C, 2] (fp64: sum, sum of squares | sum g, sum g*xhat).

    All S segment groups travel in ONE message (the reference issues one collective per segment call).
    Returns the global element count per (group, channel): equal per-rank batches, as under DistributedSampler.
    """
    dist.all_reduce(sums, group=pg)
    return count * dist.get_world_size(pg)


def max_over_ranks(ms, device=None):
    """max of a per-rank scalar (device time in ms) over all ranks."""
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def shard_batch(global_batch, world_size):
    """per-rank batch as train_adamml.py:122 (`-b` is the global batch)."""
    return int(global_batch / world_size)
