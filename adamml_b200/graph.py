"""CUDA-graph capture of a whole training step.

The engine issues ~2,100 C-ABI launches per step from Python (one per fused op, all S segments batched); at
the N=72 configuration enqueueing them takes about as long as the GPU needs to run them.  The launches are
shape-static, allocation goes through torch's caching allocator and there is no host synchronisation inside a
step, so forward + loss + backward + optimizer steps are captured ONCE into a CUDA graph (tensor-map
descriptors are baked in as kernel parameters; the graph's private memory pool keeps every address stable) and
replayed with a single launch per step — CUDA streams/graphs instead of a tracing compiler.
"""
import gc

import torch

from . import _lib


def step_guard(model, *optimizers):
    """-> callable returning a snapshot of the HOST state a captured AdaMML training step bakes into its launches:

    * the Gumbel temperature (a Python float passed to the policy kernels; `decay_temperature()` changes it every
      epoch, train_adamml.py:516),
    * the freeze / unfreeze pattern (`requires_grad` of every parameter decides which backward launches exist,
      adamml.py:111-132) and the train / eval flags (BatchNorm mode, dropout),
    * every optimizer hyper-parameter held as a Python number (lr of a non-capturable optimizer, momentum, weight
      decay); tensor-valued hyper-parameters (capturable optimizers) are read on the device and need no guard.
    """
    def snap():
        net = getattr(model, "module", model)
        pol = getattr(net, "policy_net", None)
        params = list(net.parameters())
        state = [getattr(pol, "temperature", None),
                 hash(tuple(p.requires_grad for p in params)),
                 hash(tuple(m.training for m in net.modules()))]
        for opt in optimizers:
            for g in opt.param_groups:
                state.append(tuple((k, v) for k, v in sorted(g.items())
                                   if k != "params" and isinstance(v, (int, float, bool, tuple)) and v is not None))
        return tuple(state)
    return snap


class GraphedTrainStep:
    """fn() -> tensor (e.g. the loss): reads its inputs from STATIC device tensors the caller refreshes in place
    (``static.copy_(new, non_blocking=True)``) before each replay.

    Requirements on fn (met by AdaMML.forward + torch losses + capturable optimizers): no host sync, no
    data-dependent Python control flow, optimizer state already initialised (run a few eager steps first),
    gradients set to None before capture (backward then writes fresh, static gradient tensors).

    A captured graph replays the launches with the HOST scalars they were issued with.  `guard` (see `step_guard`)
    snapshots that host state at capture time; a replay after it changed (temperature decay, LR scheduler step on a
    non-capturable optimizer, freeze_*/unfreeze_*, .train()/.eval()) raises instead of silently training with stale
    values — call `capture()` again (gradients set to None first) to pick the new state up.
    """

    def __init__(self, fn, guard=None):
        self.fn = fn
        self.guard = guard
        self._guard_state = None
        self.graph = None
        self.out = None
        self.launches = 0

    def capture(self):
        torch.cuda.synchronize()
        self.graph = None
        self.out = None
        gc.collect()
        torch.cuda.empty_cache()  # eager warm-up steps leave ~the whole working set cached in the default pool
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.out = self.fn()
        self.launches = _lib.launch_count() - n0
        self._guard_state = self.guard() if self.guard is not None else None
        torch.cuda.synchronize()
        return self

    def stale(self):
        """True when the host state baked into the captured launches has changed since capture()"""
        return self.guard is not None and self.guard() != self._guard_state

    def __call__(self):
        if self.stale():
            raise RuntimeError("GraphedTrainStep: temperature / learning rate / freeze pattern / train-eval mode "
                               "changed since capture(); the captured launches still use the old values — call "
                               "capture() again")
        self.graph.replay()
        return self.out
