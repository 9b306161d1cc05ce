"""Tensor-level wrappers over the C ABI: graph plans, row gathers, segmented reductions, fused blocks.

PyTorch is used for device memory and the current stream only; every function here requires CUDA
tensors and raises otherwise (the north star forbids a CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Optional

import torch

from . import lib as _l

D = 128


def _require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "aero_gnn_b200: the message-passing path runs only on CUDA (sm_100a) tensors; "
                f"got a tensor on {t.device} (no CPU fallback)"
            )


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _l.AERO_F32
    if t.dtype == torch.bfloat16:
        return _l.AERO_BF16
    raise RuntimeError(f"aero_gnn_b200: unsupported dtype {t.dtype} for the fused path (float32 and bfloat16 only)")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


class LaunchCounter:
    """Counts kernels launched by this library (bench.py `gpu_launches`)."""

    total = 0

    @classmethod
    def add(cls) -> None:
        cls.total += _l.load().aero_last_launch_count()

    @classmethod
    def bump(cls, n: int) -> None:
        """Launches replayed by a CUDA graph (graphs.GraphedStep)."""
        cls.total += int(n)


class _Profile:
    """Optional CUDA-event timing of the fused block launches on the launching stream (bench.py roofline)."""

    def __init__(self):
        self.enabled = False
        self.events = {}

    def reset(self, enabled: bool) -> None:
        self.enabled = enabled
        self.events = {}

    def begin(self, kind):
        if not (self.enabled and kind):
            return None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        return (kind, a, b)

    def end(self, tok) -> None:
        if tok is not None:
            kind, a, b = tok
            b.record()
            self.events.setdefault(kind, []).append((a, b))

    def summary(self):
        torch.cuda.synchronize()
        return {k: {"count": len(v), "ms_total": sum(a.elapsed_time(b) for a, b in v)} for k, v in self.events.items()}


PROFILE = _Profile()


class Fallbacks:
    """Work that left the fused sm_100a kernels for a chain of library (ATen / cuBLAS) ops ON THE GPU, counted and
    announced once per reason so that it never happens silently.  There is no CPU fallback anywhere; these are the
    shapes / activations the kernels do not cover: MLP widths other than 128 (models/mlp.py) and activations whose
    derivative is not a function of their output (SiLU, GELU, ...: processor.eager_stack)."""

    counts: dict = {}
    _warned: set = set()

    @classmethod
    def note(cls, reason: str, detail: str) -> None:
        cls.counts[reason] = cls.counts.get(reason, 0) + 1
        if reason not in cls._warned:
            cls._warned.add(reason)
            import warnings
            warnings.warn(f"aero_gnn_b200: {detail} -- running a chain of library ops on the GPU instead of the fused "
                          f"sm_100a kernels (counted in ops.Fallbacks.counts['{reason}'])", RuntimeWarning, stacklevel=3)


# ------------------------------------------------------------------------------------------------
# graph plan
# ------------------------------------------------------------------------------------------------
@dataclass
class GraphPlan:
    """Receiver-CSR + sender-CSR of one mesh (built once, cached)."""

    E: int
    N: int
    rowptr: torch.Tensor   # [N+1] int32
    perm: torch.Tensor     # [E] int32: caller edge id at CSR slot k
    src: torch.Tensor      # [E] int32 sender of CSR slot
    dst: torch.Tensor      # [E] int32 receiver of CSR slot
    sptr: torch.Tensor     # [N+1] int32
    sperm: torch.Tensor    # [E] int32 CSR slots grouped by sender
    _inv_perm: Optional[torch.Tensor] = field(default=None, repr=False)
    _inv_deg: Optional[torch.Tensor] = field(default=None, repr=False)

    @property
    def inv_perm(self) -> torch.Tensor:
        """CSR slot of caller edge id (int32)."""
        if self._inv_perm is None:
            inv = torch.empty_like(self.perm)
            inv[self.perm.long()] = torch.arange(self.E, dtype=torch.int32, device=self.perm.device)
            self._inv_perm = inv
        return self._inv_perm

    @property
    def inv_deg(self) -> torch.Tensor:
        """1 / max(in-degree, 1) as fp32 [N] (scatter_mean divisor, mgnLayer.py:144)."""
        if self._inv_deg is None:
            deg = (self.rowptr[1:] - self.rowptr[:-1]).clamp(min=1).to(torch.float32)
            self._inv_deg = 1.0 / deg
        return self._inv_deg


def build_graph_plan(edge_index: torch.Tensor, num_nodes: int) -> GraphPlan:
    _require_cuda(edge_index)
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError("edge_index must have shape [2, E]")
    lib = _l.load()
    ei = edge_index.long().contiguous()
    E, N = int(ei.size(1)), int(num_nodes)
    dev = ei.device
    i32 = dict(dtype=torch.int32, device=dev)
    rowptr = torch.empty(N + 1, **i32)
    sptr = torch.empty(N + 1, **i32)
    perm, src, dst, sperm = (torch.empty(E, **i32) for _ in range(4))
    status = torch.zeros(4, **i32)
    ws = _workspace(lib.aero_graph_plan_workspace_bytes(E, N), dev)
    with torch.cuda.device(dev):
        rc = lib.aero_graph_plan_build(_ptr(ei), E, N, _ptr(rowptr), _ptr(perm), _ptr(src), _ptr(dst), _ptr(sptr),
                                       _ptr(sperm), _ptr(status), _ptr(ws), ws.numel(), _stream())
    _l.check(rc, "aero_graph_plan_build")
    bad = int(status[0].item())
    if bad:
        raise IndexError(f"edge_index holds {bad} entries outside [0, {N})")
    return GraphPlan(E, N, rowptr, perm, src, dst, sptr, sperm)


def content_key(t: Optional[torch.Tensor]) -> tuple:
    """Cache key of a device tensor's contents: the 64-bit kernel hash plus an independent 64-bit check word (a
    position-weighted wrap-around sum computed by a different code path), so a hit on one mesh's entry by another
    mesh would need both to collide."""
    if t is None:
        return (0, 0, 0)
    tc = t.contiguous()
    raw = tc.reshape(-1).view(torch.uint8)   # (an empty [2, 0] view may carry a non-unit last stride)
    if raw.numel() % 8:
        raw = torch.cat([raw, raw.new_zeros(8 - raw.numel() % 8)])
    words = raw.view(torch.int64)
    weights = torch.arange(1, words.numel() + 1, device=words.device, dtype=torch.int64) * 0x1E3779B97F4A7C15
    check = int((words * weights).sum().item())              # int64 wrap-around arithmetic
    return (content_hash(words), check, int(words.numel()))


def content_hash(t: torch.Tensor) -> int:
    _require_cuda(t)
    lib = _l.load()
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    if nbytes % 8:
        raise ValueError("content_hash needs a multiple of 8 bytes")
    out = torch.zeros(1, dtype=torch.int64, device=t.device)
    with torch.cuda.device(t.device):
        rc = lib.aero_hash_u64(_ptr(t), nbytes, _ptr(out), _stream())
    _l.check(rc, "aero_hash_u64")
    return int(out.item())


class PlanCache:
    """Graph plans keyed by mesh connectivity.

    Fast path: a live tensor object seen before, at the same version -> no device work at all.
    Otherwise a 64-bit content hash of edge_index (one pass over 16E bytes + one 8-byte readback)
    finds the plan of a mesh seen before (DataLoader epochs revisit the same meshes).
    """

    def __init__(self, capacity: int = 64):
        self.capacity = capacity
        self._by_hash: "OrderedDict[tuple, GraphPlan]" = OrderedDict()
        self._last: dict = {}  # id(edge_index) -> (weakref, version, ptr, N, plan): every level of a hierarchy has a slot

    def get(self, edge_index: torch.Tensor, num_nodes: int) -> GraphPlan:
        last = self._last.get(id(edge_index))
        if last is not None:
            ref, ver, ptr, n, plan = last
            if ref() is edge_index and ver == edge_index._version and ptr == edge_index.data_ptr() and n == num_nodes:
                return plan
        ei = edge_index.long().contiguous()
        key = (content_key(ei), int(ei.size(1)), int(num_nodes), str(ei.device))
        plan = self._by_hash.get(key)
        if plan is None:
            plan = build_graph_plan(ei, num_nodes)
            self._by_hash[key] = plan
            while len(self._by_hash) > self.capacity:
                self._by_hash.popitem(last=False)
        else:
            self._by_hash.move_to_end(key)
        try:
            if len(self._last) >= 64:     # identities of dead tensors pile up over an epoch of fresh batches
                self._last.clear()
            self._last[id(edge_index)] = (weakref.ref(edge_index), edge_index._version, edge_index.data_ptr(), num_nodes, plan)
        except TypeError:
            pass
        return plan

    def clear(self) -> None:
        self._by_hash.clear()
        self._last.clear()


PLAN_CACHE = PlanCache()


# ------------------------------------------------------------------------------------------------
# gathers / segmented reductions
# ------------------------------------------------------------------------------------------------
def gather_rows(inp: torch.Tensor, idx: Optional[torch.Tensor], add: Optional[torch.Tensor] = None,
                n_out: Optional[int] = None) -> torch.Tensor:
    """out[i] = inp[idx[i]] (+ add[i]).  idx int32 (None = identity)."""
    _require_cuda(inp, idx, add)
    lib = _l.load()
    inp = inp.contiguous()
    n = int(idx.numel()) if idx is not None else (int(n_out) if n_out is not None else inp.size(0))
    width = inp.size(1)
    out = torch.empty((n, width), dtype=inp.dtype, device=inp.device)
    if add is not None:
        add = add.contiguous()
        assert add.shape == out.shape and add.dtype == inp.dtype
    with torch.cuda.device(inp.device):
        rc = lib.aero_gather_rows(_ptr(inp), _ptr(idx), _ptr(add), _ptr(out), n, width, dtype_code(inp), _stream())
    _l.check(rc, "aero_gather_rows")
    LaunchCounter.add()
    return out


def segment_reduce(inp: torch.Tensor, ptr: torch.Tensor, lst: Optional[torch.Tensor], n_seg: int,
                   mean: bool = False, out_dtype: Optional[torch.dtype] = None,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[n] = sum (or mean) of inp[lst[k]] for k in [ptr[n], ptr[n+1]); fixed order, fp32 accumulate.
    `out` may be a column block of a wider row-major matrix (unit column stride)."""
    _require_cuda(inp, ptr, lst)
    lib = _l.load()
    inp = inp.contiguous()
    width = inp.size(1)
    if out is None:
        out = torch.empty((n_seg, width), dtype=out_dtype or inp.dtype, device=inp.device)
    elif out.dim() != 2 or out.size(0) != n_seg or out.size(1) != width or out.stride(1) != 1 or not out.is_cuda:
        raise RuntimeError("segment_reduce: `out` must be a [n_seg, width] CUDA view with unit column stride")
    with torch.cuda.device(inp.device):
        rc = lib.aero_segment_reduce_ld(_ptr(inp), _ptr(ptr), _ptr(lst), _ptr(out), n_seg, width,
                                        out.stride(0) if n_seg > 1 else max(out.stride(0), width), dtype_code(inp),
                                        dtype_code(out), int(mean), _stream())
    _l.check(rc, "aero_segment_reduce")
    LaunchCounter.add()
    return out


def segment_bcast(g_out: torch.Tensor, seg_of_row: torch.Tensor, ptr: Optional[torch.Tensor], mean: bool) -> torch.Tensor:
    _require_cuda(g_out, seg_of_row, ptr)
    lib = _l.load()
    g_out = g_out.contiguous()
    n, width = int(seg_of_row.numel()), g_out.size(1)
    g_in = torch.empty((n, width), dtype=g_out.dtype, device=g_out.device)
    with torch.cuda.device(g_out.device):
        rc = lib.aero_segment_bcast(_ptr(g_out), _ptr(seg_of_row), _ptr(ptr), _ptr(g_in), n, width, dtype_code(g_out),
                                    int(mean), _stream())
    _l.check(rc, "aero_segment_bcast")
    LaunchCounter.add()
    return g_in


# ------------------------------------------------------------------------------------------------
# fused block
# ------------------------------------------------------------------------------------------------
def packed_floats(L: int) -> int:
    return (L + 2) * D * D + (L + 4) * D   # last D floats: gradient-only slot of the first Linear's bias


UMMA_MAX_L = 2


def choose_path(dtype: torch.dtype, act: str, L: int = 0, backward: bool = False) -> int:
    """bf16 rows run on tcgen05 when the library has the kernel; fp32 rows run the exact CUDA-core path."""
    if os.environ.get("AERO_FORCE_SIMT", "0") == "1":
        return _l.AERO_PATH_SIMT
    lib = _l.load()
    have = lib.aero_has_umma_bwd() if backward else lib.aero_has_umma()
    if dtype == torch.bfloat16 and L <= UMMA_MAX_L and have:
        return _l.AERO_PATH_UMMA
    return _l.AERO_PATH_SIMT


def keeps_hidden(keep_h0: bool, L_edge: int, L_node: int, rows_e: int, rows_n: int, K: int, device) -> bool:
    """Keep-all policy (opt-in): with kept h_0 and L = 2 blocks the forward also keeps H_1 and H_2 of both blocks
    (K steps x (E + N) x 512 B) and the backward kernels skip the recompute of the hidden layers.  Same bits either way.
    Measured on C5 (B200): edge backward 3.78 -> 3.37 ms, node backward 0.75 -> 0.71 ms, but the forward's extra tile
    stores cost as much (edge 2.18 -> 2.49 ms, node 0.41 -> 0.50 ms): 137.8 vs 137.4 ms per step for +52 GB of kept
    rows -- a wash, so the default stays the lean policy (AERO_KEEP_ACTS=h0); =all selects this one, =auto selects it
    when the extra rows take less than 40 % of the free device memory."""
    if not keep_h0 or L_edge != 2 or L_node != 2 or os.environ.get("AERO_BWD_V1", "0") == "1":
        return False
    mode = os.environ.get("AERO_KEEP_ACTS", "h0")
    if mode == "all":
        return True
    if mode != "auto":
        return False
    free, _total = torch.cuda.mem_get_info(device)
    return K * (rows_e + rows_n) * 2 * D * 2 < 0.4 * free


def keeps_h0(*paths: int) -> bool:
    """The first hidden activation is kept from the forward when every kernel of a stack runs on tcgen05
    (AERO_KEEP_H0=0 switches back to recomputing it, trading 256 B/row of memory for time)."""
    return os.environ.get("AERO_KEEP_H0", "1") != "0" and all(p == _l.AERO_PATH_UMMA for p in paths)


class PreparedBlock:
    """Device image of one block's weights for a kernel path (rebuilt when the weights change)."""

    def __init__(self, w_packed: torch.Tensor, L: int, path: int, act: str, use_ln: bool):
        _require_cuda(w_packed)
        if act not in _l.ACT_CODES:
            raise RuntimeError(
                f"aero_gnn_b200: activation '{act}' is not supported by the fused sm_100a path "
                f"(supported: {sorted(_l.ACT_CODES)})"
            )
        lib = _l.load()
        assert w_packed.dtype == torch.float32 and w_packed.numel() == packed_floats(L)
        self.L, self.path, self.act, self.use_ln = L, path, _l.ACT_CODES[act], int(use_ln)
        self.w = w_packed.contiguous()
        self.buf = _workspace(lib.aero_block_prepared_bytes(L, path), w_packed.device)
        with torch.cuda.device(w_packed.device):
            rc = lib.aero_block_prepare(_ptr(self.w), L, path, _ptr(self.buf), _stream())
        _l.check(rc, "aero_block_prepare")
        LaunchCounter.add()


def _desc(prep: PreparedBlock, main, resid, P, idx0, idx1, poff0, poff1, *, main_scale=None, rowptr=None, n_nodes=0,
          like=None):
    """`P` may be None in a backward that reads the kept h_0 rows (`like` then gives the latent dtype)."""
    d = _l.BlockDesc()
    if P is None:
        P_dtype, P_ptr, ldp = like.dtype, None, 0
        d.dtype = dtype_code(like)
    else:
        P_dtype, P_ptr, ldp = P.dtype, P.data_ptr(), P.size(1)
        d.dtype = dtype_code(P)
    d.path = prep.path
    d.L = prep.L
    d.act = prep.act
    d.use_ln = prep.use_ln
    d.main_f32 = int(main.dtype == torch.float32 and P_dtype != torch.float32)
    d.rows = main.size(0)
    d.n_nodes = n_nodes
    d.ldp = ldp
    d.poff0, d.poff1 = poff0, poff1
    d.main = main.data_ptr()
    d.main_scale = main_scale.data_ptr() if main_scale is not None else None
    d.resid = resid.data_ptr() if resid is not None else None
    d.P = P_ptr
    d.idx0 = idx0.data_ptr() if idx0 is not None else None
    d.idx1 = idx1.data_ptr() if idx1 is not None else None
    d.rowptr = rowptr.data_ptr() if rowptr is not None else None
    d.prepared = prep.buf.data_ptr()
    return d


def block_fwd(prep: PreparedBlock, main, resid, P, idx0, idx1, poff0, poff1, *, main_scale=None, rowptr=None,
              want_agg=False, kind=None, h0_out: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
              agg_out: Optional[torch.Tensor] = None, agg_clear: bool = True, rows: Optional[tuple] = None,
              main_lat_out: Optional[torch.Tensor] = None, hidden_out: Optional[tuple] = None):
    """Forward of one fused block; returns (out, agg or None).  `h0_out` ([rows,128], latent dtype, tcgen05 path
    only) receives the first hidden activation so that the backward does not have to recompute it.

    `out` / `agg_out`: caller-provided result buffers (a stack writes x' straight into the next step's extended row
    matrix; two launches over complementary row ranges share one aggregate).  `rows = (r0, r1)` runs the launch over
    that sub-range of the rows only -- every per-row tensor (main, resid, idx0, idx1, main_scale, out, h0_out) is
    addressed at row r0 + i, `rowptr` must then be the receiver CSR RELATIVE to r0 (rowptr - r0; r0 has to be a
    segment boundary), and `agg_clear=False` keeps what an earlier launch put into `agg_out`."""
    _require_cuda(main, resid, P)
    lib = _l.load()
    n_nodes = P.size(0)
    d = _desc(prep, main, resid, P, idx0, idx1, poff0, poff1, main_scale=main_scale, rowptr=rowptr, n_nodes=n_nodes)
    if out is None:
        out = torch.empty((main.size(0), D), dtype=P.dtype, device=P.device)
    elif out.shape != (main.size(0), D) or out.dtype != P.dtype or not out.is_contiguous():
        raise RuntimeError("block_fwd: `out` must be a contiguous [rows, 128] tensor of the latent dtype")
    if agg_out is not None:
        if agg_out.shape != (n_nodes, D) or agg_out.dtype != torch.float32 or not agg_out.is_contiguous():
            raise RuntimeError("block_fwd: `agg_out` must be a contiguous fp32 [n_nodes, 128] tensor")
        agg = agg_out
    else:
        agg = torch.empty((n_nodes, D), dtype=torch.float32, device=P.device) if want_agg else None
    d.out = out.data_ptr()
    d.agg = agg.data_ptr() if agg is not None else None
    d.h0 = h0_out.data_ptr() if h0_out is not None else None
    # tcgen05 path, fp32 `main` (node block): latent-dtype copy of the staged rows round(main * main_scale)
    d.main_lat = main_lat_out.data_ptr() if main_lat_out is not None else None
    if hidden_out is not None:      # keep-all policy: H_1, H_2 stored for a backward without recompute (L == 2)
        d.h_hidden[0], d.h_hidden[1] = hidden_out[0].data_ptr(), hidden_out[1].data_ptr()
    if not agg_clear:
        d.flags = _l.AERO_BLOCK_AGG_NO_CLEAR
    if rows is not None:
        r0, r1 = int(rows[0]), int(rows[1])
        if not (0 <= r0 <= r1 <= main.size(0)):
            raise RuntimeError("block_fwd: bad row range")
        if idx0 is None:
            raise RuntimeError("block_fwd: a row range needs explicit gather indices (identity is relative to r0)")
        d.rows = r1 - r0
        d.main = main.data_ptr() + r0 * D * main.element_size()
        d.out = out.data_ptr() + r0 * D * out.element_size()
        if resid is not None:
            d.resid = resid.data_ptr() + r0 * D * resid.element_size()
        if main_scale is not None:
            d.main_scale = main_scale.data_ptr() + r0 * 4
        if idx0 is not None:
            d.idx0 = idx0.data_ptr() + r0 * 4
        if idx1 is not None:
            d.idx1 = idx1.data_ptr() + r0 * 4
        if h0_out is not None:
            d.h0 = h0_out.data_ptr() + r0 * D * h0_out.element_size()
        if hidden_out is not None:
            for i in range(2):
                d.h_hidden[i] = hidden_out[i].data_ptr() + r0 * D * hidden_out[i].element_size()
    ws = _workspace(lib.aero_block_workspace_bytes(C.byref(d), 0), P.device)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    tok = PROFILE.begin(kind)
    with torch.cuda.device(P.device):
        rc = lib.aero_block_fwd(C.byref(d), _stream())
    PROFILE.end(tok)
    _l.check(rc, "aero_block_fwd")
    LaunchCounter.add()
    return out, agg


def wgrad_into(g_w: torch.Tensor, g_h0: torch.Tensor, rows_in: torch.Tensor) -> None:
    """g_w[W_main slot] = g_h0^T rows_in, the weight gradient of a block's first Linear: a plain library GEMM over the
    rows, written in fp32 straight into the packed gradient vector (no bf16 rounding of the result, no cast/copy
    launches)."""
    torch.mm(g_h0.t(), rows_in, out_dtype=torch.float32, out=g_w[: D * D].view(D, D))


def own_wgrad(dtype: torch.dtype) -> bool:
    """bf16 rows on a tcgen05 device go through aero_wgrad; AERO_WGRAD=lib keeps the library GEMMs (A/B measurements)."""
    return dtype == torch.bfloat16 and bool(_l.load().aero_has_umma()) and os.environ.get("AERO_WGRAD", "own") != "lib"


def wgrad(A: torch.Tensor, B: torch.Tensor, out: torch.Tensor, *, seg=None) -> None:
    """out[128 a, 128] (fp32, contiguous) = A[rows, 128 a]^T @ B[rows, 128] on the warp-specialised tcgen05 / TMA kernel
    (aero_wgrad); A may be a column block of a wider row matrix (unit column stride).  `seg = (dst, rowptr, n_nodes,
    seg_out)`: also write the receiver sums of A's first 128 columns -- rows in receiver-CSR order -- into
    seg_out [n_nodes, 128] (a column block of a wider matrix is fine).  bf16 rows only; fp32 rows use torch.mm."""
    _require_cuda(A, B, out)
    lib = _l.load()
    rows = int(A.size(0))
    a = A.size(1) // D
    if A.dim() == 2 and (A.stride(1) != 1 or A.stride(0) % 8 or A.data_ptr() % 16):
        A = A.contiguous()
    if not B.is_contiguous() or B.data_ptr() % 16:
        B = B.contiguous().clone() if B.data_ptr() % 16 else B.contiguous()
    if (A.dtype != torch.bfloat16 or B.dtype != torch.bfloat16 or A.dim() != 2 or A.size(1) != a * D or a not in (1, 2)
            or A.stride(1) != 1 or B.shape != (rows, D) or not B.is_contiguous() or out.dtype != torch.float32
            or out.shape != (a * D, D) or not out.is_contiguous()):
        raise RuntimeError("wgrad: A [rows, 128|256] and B [rows, 128] bf16 (unit column stride), out fp32 [128 a, 128]")
    lda = A.stride(0) if rows > 1 else max(A.stride(0), A.size(1))
    dst = rowptr = seg_out = None
    n_nodes, seg_ld = 0, 0
    if seg is not None:
        dst, rowptr, n_nodes, seg_out = seg
        if (seg_out.dim() != 2 or seg_out.size(1) != D or seg_out.stride(1) != 1 or seg_out.dtype != torch.bfloat16
                or seg_out.size(0) < n_nodes or rowptr.numel() < n_nodes + 1 or dst.numel() != rows):
            raise RuntimeError("wgrad: bad receiver-sum arguments")
        seg_ld = seg_out.stride(0) if seg_out.size(0) > 1 else max(seg_out.stride(0), D)
        if dst.data_ptr() % 16:      # the ids travel by TMA: 16-byte aligned base
            dst = dst.clone()
    ws = _workspace(lib.aero_wgrad_workspace_bytes(rows, a, int(seg is not None)), A.device)
    with torch.cuda.device(A.device):
        rc = lib.aero_wgrad(_ptr(A), lda, a, _ptr(B), rows, _ptr(out), _ptr(dst), _ptr(rowptr), int(n_nodes),
                            _ptr(seg_out), seg_ld, _ptr(ws), ws.numel(), _stream())
    _l.check(rc, "aero_wgrad")
    LaunchCounter.add()


def row_gemm(blocks, W: torch.Tensor, *, w_mn: bool, nb: int = 1, bias: Optional[torch.Tensor] = None,
             add: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[rows, 128 nb] = sum_ka blocks[ka] @ Wtile(ka) (+ bias) (+ add) on the warp-specialised tcgen05 / TMA kernel
    (aero_row_gemm).  `blocks`: 1..3 bf16 [rows,128] matrices (column blocks of wider matrices are fine); W: bf16
    [128 * len(blocks) * nb, 128] contiguous; w_mn=False: out = A W^T (nn.Linear layout, nb output blocks),
    w_mn=True: out = [A_0 | A_1 | ..] W."""
    _require_cuda(W, *blocks)
    lib = _l.load()
    na = len(blocks)
    rows = int(blocks[0].size(0))
    fixed = []
    for b in blocks:
        if b.dtype != torch.bfloat16 or b.dim() != 2 or b.size(0) != rows or b.size(1) != D:
            raise RuntimeError("row_gemm: every K-block must be bf16 [rows, 128]")
        if b.stride(1) != 1 or (rows > 1 and b.stride(0) % 8) or b.data_ptr() % 16:
            b = b.contiguous()
        fixed.append(b)
    if W.dtype != torch.bfloat16 or W.shape != (D * na * nb, D) or not W.is_contiguous() or W.data_ptr() % 16:
        raise RuntimeError("row_gemm: W must be a contiguous, 16-byte aligned bf16 [128 na nb, 128] matrix")
    if out is None:
        out = torch.empty((rows, D * nb), dtype=torch.bfloat16, device=W.device)
    if (out.dtype != torch.bfloat16 or out.shape != (rows, D * nb) or out.stride(1) != 1 or out.data_ptr() % 16
            or (rows > 1 and out.stride(0) % 8)):
        raise RuntimeError("row_gemm: out must be bf16 [rows, 128 nb] with unit column stride")
    if bias is not None and (bias.dtype != torch.bfloat16 or bias.numel() != D * nb or not bias.is_contiguous()):
        raise RuntimeError("row_gemm: bias must be a contiguous bf16 [128 nb] vector")
    if bias is not None and bias.data_ptr() % 16:
        bias = bias.clone()
    if add is not None:
        if add.dtype != torch.bfloat16 or add.shape != (rows, D * nb):
            raise RuntimeError("row_gemm: add must be bf16 [rows, 128 nb]")
        if add.stride(1) != 1 or (rows > 1 and add.stride(0) % 8) or add.data_ptr() % 16:
            add = add.contiguous()
    if rows == 0:
        return out
    ld = lambda t: t.stride(0) if rows > 1 else max(t.stride(0), t.size(1))
    ptrs = (C.c_void_p * 3)(*[b.data_ptr() for b in fixed], *([None] * (3 - na)))
    lds = (C.c_int64 * 3)(*[ld(b) for b in fixed], *([0] * (3 - na)))
    with torch.cuda.device(W.device):
        rc = lib.aero_row_gemm(ptrs, lds, na, _ptr(W), int(bool(w_mn)), nb, _ptr(bias), _ptr(add),
                               ld(add) if add is not None else 0, _ptr(out), ld(out), rows, _stream())
    _l.check(rc, "aero_row_gemm")
    LaunchCounter.add()
    return out


class ThinLinearFn(torch.autograd.Function):
    """z = x W^T + b for K = in_features <= 16 and 128 outputs (the encoders' first Linear, mgn.py:123-124) on
    aero_thin_linear_{fwd,bwd}: d(W) and d(b) come from one pass over the gradient rows."""

    @staticmethod
    def forward(ctx, x, W, b):
        _require_cuda(x, W, b)
        lib = _l.load()
        x = x if x.stride(1) == 1 else x.contiguous()
        rows, K = int(x.size(0)), int(x.size(1))
        Wc = W.detach().to(x.dtype).contiguous()
        bc = b.detach().to(x.dtype).contiguous() if b is not None else None
        out = torch.empty((rows, D), dtype=x.dtype, device=x.device)
        ldx = x.stride(0) if rows > 1 else max(x.stride(0), K)
        with torch.cuda.device(x.device):
            rc = lib.aero_thin_linear_fwd(_ptr(x), ldx, _ptr(Wc), _ptr(bc), _ptr(out), rows, K, dtype_code(x), _stream())
        _l.check(rc, "aero_thin_linear_fwd")
        LaunchCounter.add()
        ctx.save_for_backward(x, Wc)
        ctx.meta = (W.dtype, b.dtype if b is not None else None, ldx)
        return out

    @staticmethod
    def backward(ctx, g):
        x, Wc = ctx.saved_tensors
        w_dtype, b_dtype, ldx = ctx.meta
        lib = _l.load()
        rows, K = int(x.size(0)), int(x.size(1))
        g = g.contiguous().to(x.dtype)
        dwb = torch.empty((D, K + 1), dtype=torch.float32, device=x.device)
        ws = _workspace(lib.aero_thin_linear_workspace_bytes(rows, K), x.device)
        with torch.cuda.device(x.device):
            rc = lib.aero_thin_linear_bwd(_ptr(g), _ptr(x), ldx, _ptr(dwb), rows, K, dtype_code(x), _ptr(ws), ws.numel(),
                                          _stream())
        _l.check(rc, "aero_thin_linear_bwd")
        LaunchCounter.add()
        g_x = g @ Wc if ctx.needs_input_grad[0] else None      # raw features never need it
        return g_x, dwb[:, :K].to(w_dtype), (dwb[:, K].to(b_dtype) if b_dtype is not None else None)


def thin_linear_ok(x: torch.Tensor, lin) -> bool:
    return (x.is_cuda and x.dim() == 2 and x.dtype in (torch.float32, torch.bfloat16) and lin.out_features == D
            and 1 <= lin.in_features <= 16 and lin.weight.dtype == x.dtype)


def block_bwd(prep: PreparedBlock, main, P, idx0, idx1, poff0, poff1, g_out, *, g_agg=None, main_scale=None,
              has_resid_grad=False, g_main_out: Optional[torch.Tensor] = None, kind=None,
              h0: Optional[torch.Tensor] = None, n_nodes: Optional[int] = None,
              rowptr: Optional[torch.Tensor] = None, g_w_out: Optional[torch.Tensor] = None,
              main_is_lat_copy: bool = False, hidden: Optional[tuple] = None):
    """Backward of one fused block; returns (g_main, g_h0, g_w_packed[fp32], W_main slot zero).  With `h0` (the
    rows kept by block_fwd(h0_out=...)) layer 0 is not recomputed and `P` may be None.  `rowptr` (receiver CSR of
    the rows, with `g_agg`) lets the TMA-fed kernel take d(beta)'s receiver part as sum_n deg(n) g_agg[n]."""
    _require_cuda(main, P, g_out, h0)
    lib = _l.load()
    if P is None and h0 is None:
        raise RuntimeError("block_bwd needs the pre-projection P or the kept h_0 rows")
    dev, ldt = g_out.device, g_out.dtype
    if g_agg is not None:
        n_nodes = g_agg.size(0)      # rows of the receiver-gradient matrix (a partitioned mesh owns fewer than plan.N)
    elif n_nodes is None:
        n_nodes = P.size(0) if P is not None else main.size(0)
    if rowptr is not None and rowptr.numel() < n_nodes + 1:
        raise RuntimeError("block_bwd: rowptr is shorter than the receiver-gradient matrix")
    d = _desc(prep, main, None, P, idx0, idx1, poff0, poff1, main_scale=main_scale, n_nodes=n_nodes, like=g_out,
              rowptr=rowptr)
    d.h0 = h0.data_ptr() if h0 is not None else None
    if hidden is not None:          # H_1, H_2 kept by block_fwd(hidden_out=...): no recompute of the hidden layers
        d.h_hidden[0], d.h_hidden[1] = hidden[0].data_ptr(), hidden[1].data_ptr()
    rows = main.size(0)
    g_main_dtype = main.dtype
    if main_is_lat_copy:
        # `main` is the latent-dtype copy block_fwd(main_lat_out=...) made of fp32 rows (node block: the aggregate);
        # with kept h_0 rows the kernel does not read `main` at all, but the gradient w.r.t. the fp32 rows is fp32
        if h0 is None:
            raise RuntimeError("block_bwd: main_is_lat_copy needs the kept h_0 rows")
        d.main_f32, g_main_dtype = 1, torch.float32
    g_main = g_main_out if g_main_out is not None else torch.empty((rows, D), dtype=g_main_dtype, device=dev)
    g_h0 = torch.empty((rows, D), dtype=ldt, device=dev)
    # every slot except W_main is written by the kernel (W_main: by the caller's wgrad_into); `g_w_out` lets a
    # stack place the packed gradient inside one flat buffer (one all-reduce, no per-parameter copies)
    if g_w_out is not None:
        assert g_w_out.dtype == torch.float32 and g_w_out.numel() == packed_floats(prep.L) and g_w_out.is_contiguous()
        g_w = g_w_out
    else:
        g_w = torch.zeros(packed_floats(prep.L), dtype=torch.float32, device=dev)
    d.has_resid_grad = int(has_resid_grad)
    d.g_out = g_out.data_ptr()
    d.g_agg = g_agg.data_ptr() if g_agg is not None else None
    d.g_main = g_main.data_ptr()
    d.g_h0 = g_h0.data_ptr()
    d.g_w = g_w.data_ptr()
    ws = _workspace(lib.aero_block_workspace_bytes(C.byref(d), 1), dev)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    tok = PROFILE.begin(kind)
    with torch.cuda.device(dev):
        rc = lib.aero_block_bwd(C.byref(d), _stream())
    PROFILE.end(tok)
    _l.check(rc, "aero_block_bwd")
    LaunchCounter.add()
    return g_main, g_h0, g_w
