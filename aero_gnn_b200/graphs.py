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
    g = GraphedStep(step)
    for batch in loader:
        x0.copy_(batch.x); e0.copy_(batch.e)
        g()                                             # replay; parameter .grad tensors are static too
"""
from __future__ import annotations

from typing import Any, Callable, Optional

import torch

from . import ops


class GraphedStep:
    """Record `fn()` -- a step over static input tensors -- into one CUDA graph and replay it.

    `fn` must not synchronise with the host, read device values on the host, or build graph plans (`PLAN_CACHE`
    lookups hash the edge list on the device and read the hash back; do them before).  Gradients that are `None`
    when the capture starts are allocated from the graph's private pool and rewritten by every replay, as in
    PyTorch's whole-network capture recipe.
    """

    def __init__(self, fn: Callable[[], Any], warmup: int = 3, device: Optional[torch.device] = None):
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
            self.graph = torch.cuda.CUDAGraph()
            l0 = ops.LaunchCounter.total
            with torch.cuda.graph(self.graph):
                self.out = fn()
            self.launches = ops.LaunchCounter.total - l0   # launches of this library recorded in the graph

    def __call__(self):
        self.graph.replay()
        ops.LaunchCounter.bump(self.launches)
        return self.out
