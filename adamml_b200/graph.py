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


class GraphedTrainStep:
    """fn() -> tensor (e.g. the loss): reads its inputs from STATIC device tensors the caller refreshes in place
    (``static.copy_(new, non_blocking=True)``) before each replay.

    Requirements on fn (met by AdaMML.forward + torch losses + capturable optimizers): no host sync, no
    data-dependent Python control flow, optimizer state already initialised (run a few eager steps first),
    gradients set to None before capture (backward then writes fresh, static gradient tensors).
    """

    def __init__(self, fn):
        self.fn = fn
        self.graph = None
        self.out = None
        self.launches = 0

    def capture(self):
        torch.cuda.synchronize()
        gc.collect()
        torch.cuda.empty_cache()  # eager warm-up steps leave ~the whole working set cached in the default pool
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.out = self.fn()
        self.launches = _lib.launch_count() - n0
        torch.cuda.synchronize()
        return self

    def __call__(self):
        self.graph.replay()
        return self.out
