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


def allreduce_grads(params, group=None):
    """Gradient averaging of DistributedDataParallel (train_adamml.py:129) as ONE flat all-reduce: used when the
    step is captured in a CUDA graph (DDP's reducer hooks are host-driven).  Parameters without a gradient are
    skipped, exactly like unused parameters under find_unused_parameters=True."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    world = dist.get_world_size(group)
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    flat.div_(world)
    views, off = [], 0
    for g in grads:
        n = g.numel()
        views.append(flat[off:off + n].view_as(g))
        off += n
    torch._foreach_copy_(grads, views)
