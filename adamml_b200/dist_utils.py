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


class P2PStats:
    """Symmetric-memory arena for the sync-BN statistic exchange over NVLink peer memory (csrc/p2p.cu).

    Every BatchNorm layer owns one slot per pass (forward statistics / backward sums): the producing kernels write
    this rank's partial sums straight into the slot, `allreduce` launches the one-shot peer-memory reduction.
    Lanes = independent streams (backbones), each with its own flag row and epoch counter.  Requires
    torch.distributed symmetric memory (CUDA P2P over NVLink); `get()` returns None when unavailable so that the
    caller falls back to the NCCL all-reduce."""

    _instances = {}
    LANES = 16

    def __init__(self, pg, device, arena_doubles=8 << 20):
        import torch.distributed._symmetric_memory as symm
        self.pg = pg
        self.world = dist.get_world_size(pg)
        self.rank = dist.get_rank(pg)
        self.arena = symm.empty(arena_doubles, dtype=torch.float64, device=device)
        self.flags = symm.empty(self.LANES * self.world, dtype=torch.int32, device=device)
        self.arena.zero_()
        self.flags.zero_()
        self._h_arena = symm.rendezvous(self.arena, pg)
        self._h_flags = symm.rendezvous(self.flags, pg)
        self.peer_bufs = torch.tensor(list(self._h_arena.buffer_ptrs), dtype=torch.int64, device=device)
        self.peer_flags = torch.tensor(list(self._h_flags.buffer_ptrs), dtype=torch.int64, device=device)
        self.epoch = torch.zeros(self.LANES, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.cursor = 0
        self.slots = {}
        torch.cuda.synchronize(device)
        dist.barrier(pg)  # every rank's flags are zero before the first exchange

    @classmethod
    def get(cls, pg, device):
        import os
        if os.environ.get("ADAMML_B200_SYNCBN_P2P", "1") == "0":
            return None
        key = (id(pg), device.index)
        if key not in cls._instances:
            # every rank must take the SAME path (a rank on NCCL while its peers spin in the peer-memory kernel is a
            # hang): probe locally, agree with a MIN all-reduce, and only then build the arena; a failure after the
            # agreement is fatal instead of a silent per-rank fallback
            why = ""
            try:
                import torch.distributed._symmetric_memory as symm  # noqa: F401
                ok = torch.cuda.is_available() and dist.get_backend(pg) == "nccl"
                if not ok:
                    why = "needs CUDA + an NCCL process group"
            except Exception as e:  # no symmetric memory in this torch build
                ok, why = False, f"{type(e).__name__}: {e}"
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=pg)
            if int(flag.item()) == 0:
                print(f"[adamml_b200] sync-BN over peer memory unavailable on some rank ({why or 'a peer'}); "
                      "all ranks use NCCL", flush=True)
                cls._instances[key] = None
            else:
                cls._instances[key] = cls(pg, device)
        return cls._instances[key]

    @classmethod
    def check_all(cls):
        """host-side check of every arena (synchronises): call at step / epoch boundaries"""
        for inst in cls._instances.values():
            if inst is not None:
                inst.check()

    def slot(self, key, n):
        """-> (offset in doubles, float64 view [n]) of this layer's slot (allocated on first use, 256-byte aligned)"""
        if key not in self.slots:
            if self.cursor + n > self.arena.numel():
                raise RuntimeError("P2PStats arena exhausted")
            self.slots[key] = (self.cursor, n)
            self.cursor += (n + 31) // 32 * 32
        off, n0 = self.slots[key]
        assert n0 == n, "a BatchNorm layer changed its statistic size"
        return off, self.arena[off:off + n]

    def allreduce(self, off, n, lane):
        from ._lib import call
        out = torch.empty(n, dtype=torch.float64, device=self.arena.device)
        call("p2p_allreduce_f64", self.peer_bufs, self.peer_flags, off, out, n, self.world, self.rank,
             lane % self.LANES, self.epoch, self.err)
        return out

    def check(self):
        """host-side check (synchronises): raises if a peer never arrived at some exchange"""
        if int(self.err.item()):
            raise RuntimeError("sync-BN peer-memory exchange timed out (a rank fell out of step)")
