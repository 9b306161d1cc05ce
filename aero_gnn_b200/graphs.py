"""CUDA-graph capture of a whole processor step (forward + backward [+ halo exchange + gradient all-reduce]).

The fused block kernels are launched through the C ABI on `torch.cuda.current_stream()`, allocate only through
torch's caching allocator and never synchronise, so a complete step -- ~54 launches per message-passing layer --
can be recorded once and replayed with a single `cudaGraphLaunch`.  This is what keeps small meshes and the
per-rank shards of a partitioned mesh from being bound by host launch overhead (the reference pays that overhead
on every `EdgeBlock`/`NodeBlock` call, mgn.py:104-106).

Usage
-----
    plan = ops.PLAN_CACHE.get(edge_index, N)            # outside the capture: building a plan synchronises
    def step():
        x, e = run_layers(net.layers, plan, x0, e0)     # x0, e0, gx: static tensors, refilled in place
        torch.autograd.backward([x], [gx])
    g = GraphedStep(step, leaves=[x0, e0, *net.layers.parameters()])
    for batch in loader:
        x0.copy_(batch.x); e0.copy_(batch.e)
        g()                                             # replay; g.grads are the static gradient tensors
        optimizer.step()                                # then optimizer.zero_grad(set_to_none=False), or nothing
"""
from __future__ import annotations

from typing import Any, Callable, Optional, Sequence

import torch

from . import ops


class GraphedStep:
    """Record `fn()` -- a step over static input tensors -- into one CUDA graph and replay it.

    `fn` must not synchronise with the host, read device values on the host, or build graph plans (`PLAN_CACHE`
    lookups hash the edge list on the device and read the hash back; do them before).

    Gradients.  A leaf whose `.grad` is `None` when the capture starts gets its gradient allocated from the graph's
    private pool and REWRITTEN by every replay (PyTorch's whole-network capture recipe); a leaf whose `.grad` already
    exists is accumulated into, in place, by every replay.  The warm-up calls of `fn` leave gradients behind, so pass
    the leaves (parameters and differentiable inputs) as `leaves`: their `.grad` is reset to `None` after the warm-up
    and before the capture, and the static gradient tensors the replays write are kept in `self.grads` (same order).
    After the capture do not call `zero_grad(set_to_none=True)` on those leaves: that drops the static tensors while
    the graph keeps writing into them.  Use `optimizer.zero_grad(set_to_none=False)` or nothing at all (every replay
    overwrites).  Without `leaves`, `fn` itself has to set the gradients to `None` (as bench.py's step does); a
    capture that starts with gradients present raises.
    """

    def __init__(self, fn: Callable[[], Any], warmup: int = 3, device: Optional[torch.device] = None,
                 leaves: Optional[Sequence[torch.Tensor]] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedStep needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        if ops.PROFILE.enabled:
            raise RuntimeError("GraphedStep: disable ops.PROFILE first (timing events cannot be captured)")
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(max(warmup, 1)):     # lazy initialisation (function attributes, NCCL channels, cuBLAS
                    fn()                            # workspaces) must happen before the capture
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.leaves = list(leaves) if leaves is not None else []
            for t in self.leaves:                   # the warm-up's gradients must not be accumulated into by replays
                t.grad = None
            self.graph = torch.cuda.CUDAGraph()
            l0 = ops.LaunchCounter.total
            with torch.cuda.graph(self.graph):
                self.out = fn()
            self.launches = ops.LaunchCounter.total - l0   # launches of this library recorded in the graph
            self.grads = [t.grad for t in self.leaves]     # static: rewritten by every replay

    def __call__(self):
        self.graph.replay()
        ops.LaunchCounter.bump(self.launches)
        return self.out
