"""Bistride pooling / unpooling between mesh levels (reference models/bsms_mgn.py:217-306).

Index construction (x-sorted rank//stride assignment, coarse edge keys, unique + inverse) is integer
work done once per mesh by the kernels behind aero_stride_pool_plan / aero_coarsen_edges and cached;
it is bit-exact against the reference.  The floating-point parts (mean-pool of node latents,
positions and edge latents; unpool gather + skip add) are deterministic segmented kernels with
autograd wrappers.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from dataclasses import dataclass
from typing import Optional

import torch

from . import lib as _l
from . import ops
from .ops import _ptr, _stream, _workspace


# ------------------------------------------------------------------------------------------------
# raw index kernels
# ------------------------------------------------------------------------------------------------
def stride_pool_assign(batch: torch.Tensor, posx: Optional[torch.Tensor], stride: int):
    """(fine_to_coarse int64 [N], coarse_batch int64 [Nc]) exactly as bsms_mgn.py:231-262."""
    ops._require_cuda(batch, posx)
    lib = _l.load()
    batch = batch.long().contiguous()
    N = int(batch.numel())
    dev = batch.device
    px = posx.to(torch.float64).contiguous() if posx is not None else None
    f2c = torch.empty(N, dtype=torch.int64, device=dev)
    cb = torch.empty(N, dtype=torch.int64, device=dev)
    counts = torch.zeros(2, dtype=torch.int64, device=dev)
    ws = _workspace(lib.aero_stride_pool_workspace_bytes(N), dev)
    with torch.cuda.device(dev):
        rc = lib.aero_stride_pool_plan(_ptr(batch), _ptr(px), N, int(stride), _ptr(f2c), _ptr(cb), _ptr(counts),
                                       _ptr(ws), ws.numel(), _stream())
    _l.check(rc, "aero_stride_pool_plan")
    ops.LaunchCounter.add()
    nc = int(counts[0].item())
    return f2c, cb[:nc].clone()


def coarsen_edges(edge_index: torch.Tensor, f2c: torch.Tensor, n_coarse: int):
    """(coarse_edge_index [2,Ec] int64, inverse [E] int64, gptr [Ec+1] int32, glist [E] int32), bsms_mgn.py:274-288."""
    ops._require_cuda(edge_index, f2c)
    lib = _l.load()
    ei = edge_index.long().contiguous()
    E = int(ei.size(1))
    dev = ei.device
    cei = torch.empty((2, max(E, 1)), dtype=torch.int64, device=dev)
    inverse = torch.empty(E, dtype=torch.int64, device=dev)
    gptr = torch.empty(E + 1, dtype=torch.int32, device=dev)
    glist = torch.empty(E, dtype=torch.int32, device=dev)
    counts = torch.zeros(2, dtype=torch.int64, device=dev)
    ws = _workspace(lib.aero_coarsen_edges_workspace_bytes(E), dev)
    with torch.cuda.device(dev):
        rc = lib.aero_coarsen_edges(_ptr(ei), E, _ptr(f2c), int(n_coarse), _ptr(cei), _ptr(inverse), _ptr(gptr),
                                    _ptr(glist), _ptr(counts), _ptr(ws), ws.numel(), _stream())
    _l.check(rc, "aero_coarsen_edges")
    ops.LaunchCounter.add()
    ec = int(counts[0].item())
    return cei[:, :ec].contiguous(), inverse, gptr[: ec + 1].contiguous(), glist


def group_lists(group_of: torch.Tensor, n_groups: int):
    """(gptr [n_groups+1] int32, glist [n] int32 members ascending, group32 [n] int32)."""
    ops._require_cuda(group_of)
    lib = _l.load()
    g = group_of.long().contiguous()
    n = int(g.numel())
    dev = g.device
    gptr = torch.empty(n_groups + 1, dtype=torch.int32, device=dev)
    glist = torch.empty(n, dtype=torch.int32, device=dev)
    g32 = torch.empty(n, dtype=torch.int32, device=dev)
    ws = _workspace(lib.aero_group_lists_workspace_bytes(n), dev)
    with torch.cuda.device(dev):
        rc = lib.aero_group_lists(_ptr(g), n, int(n_groups), _ptr(gptr), _ptr(glist), _ptr(g32), _ptr(ws), ws.numel(),
                                  _stream())
    _l.check(rc, "aero_group_lists")
    ops.LaunchCounter.add()
    return gptr, glist, g32


# ------------------------------------------------------------------------------------------------
# autograd wrappers for the floating-point parts
# ------------------------------------------------------------------------------------------------
class SegmentReduceFn(torch.autograd.Function):
    """out[g] = sum|mean of inp[glist[gptr[g]:gptr[g+1]]] (torch_scatter.scatter_add / scatter_mean with a
    precomputed, ascending member list -- the order CPU scatter_add_ accumulates in)."""

    @staticmethod
    def forward(ctx, inp, gptr, glist, group_of_row, n_groups: int, mean: bool):
        ctx.meta = (gptr, group_of_row, mean)
        return ops.segment_reduce(inp, gptr, glist, n_groups, mean=mean)

    @staticmethod
    def backward(ctx, g):
        gptr, group_of_row, mean = ctx.meta
        return ops.segment_bcast(g.contiguous(), group_of_row, gptr, mean), None, None, None, None, None


class UnpoolAddFn(torch.autograd.Function):
    """fine = coarse[assignment] + skip (bsms_mgn.py:199-200, :303-306)."""

    @staticmethod
    def forward(ctx, coarse, skip, assign32, gptr, glist):
        ctx.meta = (gptr, glist, coarse.size(0))
        return ops.gather_rows(coarse, assign32, add=skip)

    @staticmethod
    def backward(ctx, g):
        gptr, glist, nc = ctx.meta
        g = g.contiguous()
        return ops.segment_reduce(g, gptr, glist, nc, mean=False), g, None, None, None


# ------------------------------------------------------------------------------------------------
# one pooling level, cached per mesh
# ------------------------------------------------------------------------------------------------
@dataclass
class PoolLevel:
    n_fine: int
    n_coarse: int
    fine_to_coarse: torch.Tensor      # int64 [N]  (the reference's `assignment`)
    f2c32: torch.Tensor               # int32 [N]
    node_gptr: torch.Tensor           # int32 [Nc+1]
    node_glist: torch.Tensor          # int32 [N]
    coarse_batch: torch.Tensor        # int64 [Nc]
    coarse_edge_index: torch.Tensor   # int64 [2, Ec], sorted by (sender, receiver)
    inverse: torch.Tensor             # int64 [E]: coarse edge of each fine edge (caller order)
    edge_gptr: torch.Tensor           # int32 [Ec+1]
    edge_glist: torch.Tensor          # int32 [E] fine caller edge ids grouped by coarse edge


def build_pool_level(edge_index, batch, pos, stride: int) -> PoolLevel:
    posx = pos[:, 0] if pos is not None else None
    f2c, cb = stride_pool_assign(batch, posx, stride)
    nc = int(cb.numel())
    node_gptr, node_glist, f2c32 = group_lists(f2c, nc)
    cei, inverse, egptr, eglist = coarsen_edges(edge_index, f2c, nc)
    return PoolLevel(int(batch.numel()), nc, f2c, f2c32, node_gptr, node_glist, cb, cei, inverse, egptr, eglist)


def _hash_any(t: Optional[torch.Tensor]) -> int:
    if t is None:
        return 0
    b = t.contiguous().view(torch.uint8).reshape(-1)
    pad = (-b.numel()) % 8
    if pad:
        b = torch.cat([b, b.new_zeros(pad)])
    return ops.content_hash(b)


def _ident(*tensors):
    """Identity of live tensors at their current version (no device work): (weakrefs, fingerprint)."""
    import weakref
    refs, fp = [], []
    for t in tensors:
        if t is None:
            refs.append(None)
            fp.append(None)
        else:
            try:
                refs.append(weakref.ref(t))
            except TypeError:
                return None, None
            fp.append((t.data_ptr(), t._version, tuple(t.shape), t.dtype))
    return refs, tuple(fp)


def _same(last, tensors, extra) -> bool:
    if last is None:
        return False
    refs, fp, ex, _ = last
    if ex != extra:
        return False
    for r, t in zip(refs, tensors):
        if (r is None) != (t is None) or (r is not None and r() is not t):
            return False
    return _ident(*tensors)[1] == fp


class PoolCache:
    """Pool levels keyed by content (kernel hash + independent check word, ops.content_key).  Fast path: the same
    live tensor objects at the same version -> no device work and no host sync at all (a hierarchy is looked up on
    every forward; three hashes with three read-backs per level would also make the forward uncapturable)."""

    def __init__(self, capacity: int = 64):
        self.capacity = capacity
        self._d: "OrderedDict[tuple, PoolLevel]" = OrderedDict()
        self._last: dict = {}

    def get(self, edge_index, batch, pos, stride: int) -> PoolLevel:
        slot = (int(stride), int(batch.numel()), int(edge_index.shape[1]))
        last = self._last.get(slot)
        if _same(last, (edge_index, batch, pos), int(stride)):
            return last[3]
        key = (ops.content_key(edge_index), ops.content_key(batch), ops.content_key(pos), int(stride),
               tuple(edge_index.shape), int(batch.numel()), str(edge_index.device))
        lvl = self._get_slow(key, edge_index, batch, pos, stride)
        refs, fp = _ident(edge_index, batch, pos)
        if refs is not None:
            if len(self._last) > 32:
                self._last.clear()
            self._last[slot] = (refs, fp, int(stride), lvl)
        return lvl

    def _get_slow(self, key, edge_index, batch, pos, stride: int) -> PoolLevel:
        lvl = self._d.get(key)
        if lvl is None:
            lvl = build_pool_level(edge_index, batch, pos, stride)
            self._d[key] = lvl
            while len(self._d) > self.capacity:
                self._d.popitem(last=False)
        else:
            self._d.move_to_end(key)
        return lvl

    def clear(self):
        self._d.clear()
        self._last.clear()


POOL_CACHE = PoolCache()
