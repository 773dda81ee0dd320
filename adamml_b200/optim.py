"""Train-step tail on the device in a handful of launches.

The reference's iteration ends with CrossEntropy + compute_policy_loss + two torch optimizers looping over ~650
parameter tensors each (utils/utils.py:362-400, train_adamml.py:250-257).  Here:

* `loss_tail(logits, target, selection, cost_weights, gammas, use_policy)` — cross-entropy + the 'blockdrop' policy loss
  (utils/utils.py:166-184, including its [N] x [N,1] broadcast) and BOTH gradients in one kernel;
* `FusedSGD` / `FusedAdam` — drop-in `torch.optim.Optimizer`s with torch's update rules (SGD: momentum, dampening 0,
  L2 weight decay; Adam: bias correction, eps outside the root, L2 weight decay, no amsgrad) that update every tensor
  of a parameter group in ONE multi-tensor launch.  The Adam step counter lives on the device, the pointer table is
  copied from pinned memory, so `step()` captures into a CUDA graph;
* `clip_grad_norm_(parameters, max_norm)` — `torch.nn.utils.clip_grad_norm_` (L2; utils/utils.py:390-391, the
  reference's `--clip_gradient`) over all gradient tensors as two multi-tensor launches, the coefficient computed on the
  device: no host synchronisation, capturable.
"""
import torch

from ._lib import call, lib


class _LossTail(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, selection, target, cost_weights, gamma, use_policy):
        logits = logits.contiguous().float()
        selection = selection.contiguous().float()
        N, C = logits.shape
        _, S, M = selection.shape
        loss = torch.empty((), device=logits.device, dtype=torch.float32)
        dlogits = torch.empty_like(logits)
        dsel = torch.empty_like(selection)
        call("loss_tail", logits, target.contiguous(), selection, cost_weights, float(gamma), int(use_policy), N, C, S, M,
             loss, dlogits, dsel)
        ctx.save_for_backward(dlogits, dsel)
        return loss

    @staticmethod
    def backward(ctx, g):
        dlogits, dsel = ctx.saved_tensors
        return g * dlogits, g * dsel, None, None, None, None


def loss_tail(logits, target, selection, cost_weights, gammas, use_policy=True):
    """loss = CE(logits, target) [+ blockdrop policy loss] as in the step body of train_adamml()
    (utils/utils.py:362-382).  selection: [N, S, M] decisions; cost_weights: fp32 [M] device tensor."""
    if target.dtype != torch.int64:
        target = target.long()
    cw = cost_weights.contiguous().float() if cost_weights is not None else None
    return _LossTail.apply(logits, selection, target, cw, float(gammas), bool(use_policy))


class _FusedOptimizer(torch.optim.Optimizer):
    N_STATE = 1  # state tensors per parameter

    def _tables(self, group, params, states):
        """device tables of one launch: [2 + N_STATE][n] pointers, sizes, chunk -> (tensor, offset).  Rebuilt only when
        an address changes (eager steps with set_to_none re-allocate the gradients; inside a captured graph every
        address is stable).  Each key owns its pinned staging buffer: a captured H2D copy re-reads it on replay."""
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in params)
        cache = group.setdefault("_adamml_tables", {})
        if key in cache:
            return cache[key]
        if len(cache) > 4:   # eager mode: gradients move every step, keep the cache small
            cache.clear()
        n = len(params)
        dev = params[0].device
        rows = [[p.data_ptr() for p in params], [p.grad.data_ptr() for p in params]]
        for k in range(self.N_STATE):
            rows.append([s[k].data_ptr() for s in states])
        host = torch.tensor(rows, dtype=torch.int64).pin_memory()
        table = torch.empty_like(host, device=dev)
        table.copy_(host, non_blocking=True)
        skey = tuple(p.numel() for p in params)
        geo = group.setdefault("_adamml_geo", {})
        if skey not in geo:
            chunk = int(lib().cdll.adamml_opt_chunk())
            ct, cs = [], []
            for i, numel in enumerate(skey):
                for off in range(0, numel, chunk):
                    ct.append(i)
                    cs.append(off)
            geo[skey] = (torch.tensor(skey, dtype=torch.int64, device=dev), torch.tensor(ct, dtype=torch.int32, device=dev),
                         torch.tensor(cs, dtype=torch.int64, device=dev), len(ct))
        cache[key] = (table, host, n) + geo[skey]
        return cache[key]

    def _live(self, group):
        params, states = [], []
        for p in group["params"]:
            if p.grad is None:
                continue
            if (p.grad.dtype != torch.float32 or p.dtype != torch.float32 or not p.is_contiguous()
                    or not p.grad.is_contiguous()):
                raise TypeError("fused optimizers take contiguous fp32 parameters and gradients")
            st = self.state[p]
            if "bufs" not in st:
                st["bufs"] = [torch.zeros_like(p, memory_format=torch.preserve_format) for _ in range(self.N_STATE)]
            params.append(p)
            states.append(st["bufs"])
        return params, states


class FusedSGD(_FusedOptimizer):
    """torch.optim.SGD(params, lr, momentum, weight_decay) semantics (dampening 0, nesterov False)."""
    N_STATE = 1

    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        for group in self.param_groups:
            params, states = self._live(group)
            if not params:
                continue
            table, _, n, sizes, ct, cs, nchunks = self._tables(group, params, states)
            call("sgd_multi", table, sizes, ct, cs, n, nchunks, float(group["lr"]), float(group["momentum"]),
                 float(group["weight_decay"]))


class FusedAdam(_FusedOptimizer):
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) semantics (amsgrad False)."""
    N_STATE = 2

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        for group in self.param_groups:
            params, states = self._live(group)
            if not params:
                continue
            if "_adamml_step" not in group:
                group["_adamml_step"] = torch.zeros(1, dtype=torch.int64, device=params[0].device)
            table, _, n, sizes, ct, cs, nchunks = self._tables(group, params, states)
            b1, b2 = group["betas"]
            call("adam_multi", table, sizes, ct, cs, n, nchunks, float(group["lr"]), float(b1), float(b2),
                 float(group["eps"]), float(group["weight_decay"]), group["_adamml_step"])


_CLIP_CACHE = {}
_CLIP_GEO = {}


@torch.no_grad()
def clip_grad_norm_(parameters, max_norm):
    """torch.nn.utils.clip_grad_norm_(parameters, max_norm, norm_type=2.0) for contiguous fp32 CUDA gradients
    (utils/utils.py:390-391).  -> total norm BEFORE clipping as a 0-dim device tensor (torch returns the same; reading it
    is the caller's synchronisation, the clipping itself needs none)."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.zeros((), dtype=torch.float32)
    for g in grads:
        if g.dtype != torch.float32 or not g.is_contiguous() or not g.is_cuda:
            raise TypeError("fused clip_grad_norm_ takes contiguous fp32 CUDA gradients")
    dev = grads[0].device
    # chunk geometry: keyed by the tensor sizes, built once (an eager warm-up step) -- inside a graph capture only the
    # pinned-memory address table below may be (re)built, exactly as in _FusedOptimizer._tables
    skey = (dev.index,) + tuple(g.numel() for g in grads)
    geo = _CLIP_GEO.get(skey)
    if geo is None:
        chunk = int(lib().cdll.adamml_opt_chunk())
        ct, cs = [], []
        for i, n in enumerate(skey[1:]):
            for off in range(0, n, chunk):
                ct.append(i)
                cs.append(off)
        geo = (torch.tensor(skey[1:], dtype=torch.int64, device=dev), torch.tensor(ct, dtype=torch.int32, device=dev),
               torch.tensor(cs, dtype=torch.int64, device=dev), len(ct), torch.empty(1, dtype=torch.float64, device=dev))
        _CLIP_GEO[skey] = geo
    key = tuple(g.data_ptr() for g in grads)
    ent = _CLIP_CACHE.get(key)
    if ent is None:
        if len(_CLIP_CACHE) > 4:  # eager steps with set_to_none re-allocate the gradients every step
            _CLIP_CACHE.clear()
        host = torch.tensor(key, dtype=torch.int64).pin_memory()
        table = torch.empty_like(host, device=dev)
        table.copy_(host, non_blocking=True)
        ent = (table, host) + geo
        _CLIP_CACHE[key] = ent
    table, _, sizes, ct, cs, nchunks, scratch = ent
    total = torch.empty((), dtype=torch.float32, device=dev)
    call("clip_grad_norm_multi", table, sizes, ct, cs, len(grads), nchunks, float(max_norm), scratch, total)
    return total
