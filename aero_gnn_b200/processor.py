"""MGN processor stack: K message-passing steps on one mesh as a single autograd node.

Follows models/mgnLayer.py:177-213 (one step) and the layer loops of models/mgn.py:127-128,
bsms_mgn.py:156-159/188-191/208-211 of the reference.  Per step (edge latents kept in
receiver-CSR order, see ops.GraphPlan):

    P      = x @ [W_s; W_d; W_nx]^T + [0; b_e0; b_n0]           node pre-projection (sum trick,
                                                                mgnLayer.py:97-103), plain GEMM
    e'     = e + LN(MLP_e(e W_e^T + P_s[src] + P_d[dst]))       fused edge block  (mgnLayer.py:192,205)
    agg    = segment_sum(e', receiver)                          fused into the edge block (:146)
    x'     = x + LN(MLP_n(agg W_na^T + P_n))                    fused node block  (mgnLayer.py:208,211)

Backward recomputes each block from the saved layer inputs (x, e) and the fp32 aggregate; no MLP
activation is kept (reference autograd keeps all of them).  Gradients of gathered node rows are
reduced by receiver-CSR / sender-CSR segmented sums -- no atomics anywhere.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import ops

D = ops.D


@dataclass
class StepWeights:
    """Packed parameters of one processor step (all tensors differentiable functions of the module params)."""

    w_edge: torch.Tensor   # fp32 [packed_floats(L_e)]  W_main = W_e
    w_node: torch.Tensor   # fp32 [packed_floats(L_n)]  W_main = W_n[:, D:2D] (aggregate half)
    w_proj: torch.Tensor   # [3D, D] latent dtype: [W_s; W_d; W_n[:, :D]]
    b_proj: torch.Tensor   # [3D]    latent dtype: [0; b_e0; b_n0]


@dataclass
class StackConfig:
    L_edge: int
    L_node: int
    act_edge: str
    act_node: str
    use_ln: bool = True
    mean: bool = False      # aggregation == 'mean' (mgnLayer.py:143-144)


def pack_block(w_main, hidden: Sequence, w_out, b_out, gamma, beta) -> torch.Tensor:
    """[W_main | W_1..W_L | W_out | b_1..b_L | b_out | gamma | beta | 0] as one fp32 vector."""
    parts = [w_main.reshape(-1)] + [w.reshape(-1) for (w, _) in hidden] + [w_out.reshape(-1)]
    parts += [b.reshape(-1) for (_, b) in hidden] + [b_out.reshape(-1), gamma.reshape(-1), beta.reshape(-1)]
    parts.append(torch.zeros_like(b_out.reshape(-1)))   # gradient-only slot (first Linear's bias lives in b_proj)
    if all(p.dtype == parts[0].dtype for p in parts):
        return torch.cat(parts).float()                 # one widening cast (exact) instead of one per parameter
    return torch.cat([p.float() for p in parts])


# ------------------------------------------------------------------------------------------------
# one-launch packing of a step's parameters (and one-launch scatter of their gradients)
# ------------------------------------------------------------------------------------------------
def _locate(t: torch.Tensor):
    """(base tensor, element offset inside it, rows, cols, row stride) of a parameter or of a column-slice view of
    one (e.g. W0[:, :128]); the base is what autograd sees, so no Slice/Cat backward nodes are created."""
    base = t._base if t._base is not None else t
    if t.dim() == 2:
        rows, cols, ld = t.size(0), t.size(1), t.stride(0)
        ok = t.stride(1) == 1
    else:
        rows, cols, ld = 1, t.numel(), t.numel()
        ok = t.dim() == 1 and (t.numel() <= 1 or t.stride(0) == 1)
    if not (ok and base.is_contiguous()):
        raise RuntimeError("pack_step: parameters must be contiguous (column slices of a contiguous matrix are fine)")
    return base, t.storage_offset() - base.storage_offset(), rows, cols, ld


class PackSpec:
    """Where every parameter piece of a step goes: outs = [(numel, dtype)], segs = [(base index | None, element offset in
    the base, rows, cols, base row stride, out index, element offset in the out, out row stride)]."""

    def __init__(self):
        self.bases, self.outs, self.segs = [], [], []

    def out(self, numel: int, dtype) -> int:
        self.outs.append((int(numel), dtype))
        return len(self.outs) - 1

    def put(self, t: Optional[torch.Tensor], out: int, off: int, rows: int = 1, cols: int = D, out_ld: Optional[int] = None):
        if t is None:                                   # zero fill
            self.segs.append((None, 0, rows, cols, cols, out, off, out_ld or cols))
            return
        base, boff, r, c, ld = _locate(t)
        if r * c != rows * cols:
            raise RuntimeError(f"pack_step: a parameter piece has {r}x{c} elements, expected {rows}x{cols}")
        for i, b in enumerate(self.bases):
            if b is base:
                break
        else:
            self.bases.append(base)
            i = len(self.bases) - 1
        self.segs.append((i, boff, r, c, ld, out, off, out_ld or c))


def _multi_copy(segs) -> None:
    lib = ops._l.load()
    for i in range(0, len(segs), ops._l.MAX_COPY_SEGS):
        chunk = segs[i: i + ops._l.MAX_COPY_SEGS]
        arr = (ops._l.CopySeg * len(chunk))(*chunk)
        rc = lib.aero_multi_copy(arr, len(chunk), ops._stream())
        ops._l.check(rc, "aero_multi_copy")
        ops.LaunchCounter.add()


class PackStepFn(torch.autograd.Function):
    """apply(spec, *spec.bases) -> one tensor per spec.outs entry, filled by ONE kernel launch; the backward scatters
    the gradients of those tensors into one gradient buffer per base with ONE launch (replaces the torch.cat /
    dtype-cast / slice chains and their autograd nodes, ~15 small launches per processor step)."""

    @staticmethod
    def forward(ctx, spec: PackSpec, *bases):
        dev = bases[0].device
        ops._require_cuda(*bases)
        outs = [torch.empty(n, dtype=dt, device=dev) for n, dt in spec.outs]
        segs = []
        for bi, boff, r, c, ld, oi, ooff, old in spec.segs:
            o = outs[oi]
            src = None if bi is None else bases[bi].data_ptr() + boff * bases[bi].element_size()
            segs.append(ops._l.CopySeg(src, o.data_ptr() + ooff * o.element_size(), r, c, ld, old,
                                       ops.dtype_code(bases[bi]) if bi is not None else ops.dtype_code(o),
                                       ops.dtype_code(o)))
        with torch.cuda.device(dev):
            _multi_copy(segs)
        ctx.spec = spec
        ctx.meta = [(b.shape, b.dtype) for b in bases]
        ctx.dev = dev
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        spec, dev = ctx.spec, ctx.dev
        gouts = [None if g is None else g.contiguous() for g in gouts]
        covered = [0] * len(ctx.meta)
        for bi, _, r, c, *_rest in spec.segs:
            if bi is not None:
                covered[bi] += r * c
        grads = []
        for i, (shape, dt) in enumerate(ctx.meta):
            if not ctx.needs_input_grad[1 + i]:
                grads.append(None)
                continue
            full = covered[i] == int(torch.Size(shape).numel())
            grads.append((torch.empty if full else torch.zeros)(shape, dtype=dt, device=dev))
        segs = []
        for bi, boff, r, c, ld, oi, ooff, old in spec.segs:
            if bi is None or grads[bi] is None:
                continue
            g, go = grads[bi], gouts[oi]
            src = None if go is None else go.data_ptr() + ooff * go.element_size()
            segs.append(ops._l.CopySeg(src, g.data_ptr() + boff * g.element_size(), r, c, old, ld,
                                       ops.dtype_code(go) if go is not None else ops.dtype_code(g), ops.dtype_code(g)))
        with torch.cuda.device(dev):
            _multi_copy(segs)
        return (None, *grads)


def pack_step_spec(edge, node, proj_w: Sequence, proj_b: Sequence, dtype: torch.dtype) -> PackSpec:
    """Segment list that packs one processor step from the reference-named parameters.  `edge` / `node` =
    (w_main, hidden [(W, b)...], w_out, b_out, gamma, beta); proj_w = the [D, D] blocks of the node pre-projection
    (sender part, receiver part, node-block part); proj_b = their biases (None = zeros)."""
    spec = PackSpec()
    outs = []
    for (w_main, hidden, w_out, b_out, gamma, beta) in (edge, node):
        L = len(hidden)
        o = spec.out(ops.packed_floats(L), torch.float32)
        outs.append(o)
        spec.put(w_main, o, 0, D, D)
        for l, (w, _) in enumerate(hidden):
            spec.put(w, o, (1 + l) * D * D, D, D)
        spec.put(w_out, o, (1 + L) * D * D, D, D)
        voff = (2 + L) * D * D
        for l, (_, b) in enumerate(hidden):
            spec.put(b, o, voff + l * D)
        spec.put(b_out, o, voff + L * D)
        spec.put(gamma, o, voff + (L + 1) * D)
        spec.put(beta, o, voff + (L + 2) * D)
        spec.put(None, o, voff + (L + 3) * D)            # gradient-only slot of the first Linear's bias
    k = len(proj_w)
    ow, ob = spec.out(k * D * D, dtype), spec.out(k * D, dtype)
    for i, (w, b) in enumerate(zip(proj_w, proj_b)):
        spec.put(w, ow, i * D * D, D, D)
        spec.put(b, ob, i * D)
    spec.dtype = dtype
    return spec


def apply_pack_spec(spec: PackSpec) -> StepWeights:
    w_edge, w_node, w_proj, b_proj = PackStepFn.apply(spec, *spec.bases)
    return StepWeights(w_edge, w_node, w_proj.view(-1, D), b_proj)


def pack_step(edge, node, proj_w: Sequence, proj_b: Sequence, dtype: torch.dtype) -> StepWeights:
    return apply_pack_spec(pack_step_spec(edge, node, proj_w, proj_b, dtype))


def cached_pack_step(module: torch.nn.Module, dtype: torch.dtype, build) -> StepWeights:
    """pack_step with the segment list remembered on `module`: it depends only on which Parameter objects the module
    holds and on their layout, not on their values or addresses (those are read at launch), so the per-step host work
    is one comparison of the parameter list instead of ~25 view constructions.  `build()` -> the pack_step arguments."""
    params = tuple(module.parameters())
    hit = module.__dict__.get("_aero_pack_cache")
    if hit is not None:
        key, spec = hit
        if spec.dtype == dtype and len(key) == len(params) and all(
                a is b and sh == b.shape and dt == b.dtype and dv == b.device for (a, sh, dt, dv), b in zip(key, params)):
            return apply_pack_spec(spec)
    spec = pack_step_spec(*build(), dtype)
    module.__dict__["_aero_pack_cache"] = (tuple((p, p.shape, p.dtype, p.device) for p in params), spec)
    return apply_pack_spec(spec)


# Cross-rank reduction of the stack's parameter gradients (data-parallel training of batched small meshes,
# train.py:50-51; the receiver-block partition passes its own).  `fn(flat_fp32_span) -> work | None` starts an
# all-reduce (async_op=True returns a work handle that is waited for before the gradients are handed to autograd);
# it is called once per bucket of `bucket_layers` processor steps, as soon as the backward has produced them, so
# the reduction of step k's gradients travels under the backward kernels of steps < k.
_GRAD_REDUCE = {"fn": None, "bucket_layers": 5}


def set_grad_reduce(fn, bucket_layers: int = 5) -> None:
    _GRAD_REDUCE["fn"], _GRAD_REDUCE["bucket_layers"] = fn, max(int(bucket_layers), 1)


class GradSink:
    """Flat fp32 buffer for the parameter gradients of a K-step stack, per step contiguous:
        [ w_edge | w_node | w_proj [3D, D] | b_proj [3D] ] x K
    The block kernels and the weight-gradient GEMMs write their results straight into its slices; a bucket of steps
    is summed across ranks with ONE all-reduce of one contiguous span while earlier steps are still in their backward;
    the projection parts are converted to the latent dtype with one launch for all steps (instead of ~25 small cast /
    cat / copy launches per step)."""

    def __init__(self, K: int, L_edge: int, L_node: int, device, reduce=None, bucket_layers: Optional[int] = None):
        self.K, self.pe, self.pn = K, ops.packed_floats(L_edge), ops.packed_floats(L_node)
        self.blk = self.pe + self.pn
        self.pp = 3 * D * D + 3 * D
        self.per = self.blk + self.pp
        self.flat = torch.empty(K * self.per, dtype=torch.float32, device=device)
        self.rows = self.flat.view(K, self.per)
        self.reduce = reduce if reduce is not None else _GRAD_REDUCE["fn"]
        self.bucket = max(int(bucket_layers or _GRAD_REDUCE["bucket_layers"]), 1)
        self.works, self.hi = [], K          # steps [hi, K) have been handed to the reduction

    def w_edge(self, k: int) -> torch.Tensor:
        return self.rows[k, : self.pe]

    def w_node(self, k: int) -> torch.Tensor:
        return self.rows[k, self.pe: self.blk]

    def w_proj(self, k: int) -> torch.Tensor:
        return self.rows[k, self.blk: self.blk + 3 * D * D].view(3 * D, D)

    def _flush(self, lo: int) -> None:
        """Steps [lo, hi) are complete: fill their b_proj slots (= column sums of g_h0, which the block kernels leave
        in the bias0 slot of their packed gradient) and start their cross-rank sum."""
        r = self.rows[lo: self.hi]
        bp = r[:, self.blk + 3 * D * D:].view(-1, 3, D)
        b0e, b0n = r[:, self.pe - D: self.pe], r[:, self.blk - D: self.blk]
        bp[:, 0].copy_(b0e)      # gradient of the (zero) sender-part bias: same column sums, unused by the caller
        bp[:, 1].copy_(b0e)
        bp[:, 2].copy_(b0n)
        if self.reduce is not None:
            w = self.reduce(self.flat[lo * self.per: self.hi * self.per])
            if w is not None:
                self.works.append(w)
        self.hi = lo

    def step_done(self, k: int) -> None:
        """The backward has written every gradient of step k (steps complete in descending order)."""
        if self.reduce is not None and self.hi - k >= self.bucket:
            self._flush(k)

    def finish(self, dtype: torch.dtype):
        """-> per step (g_w_edge, g_w_node, g_w_proj, g_b_proj); waits for the outstanding reductions."""
        if self.hi > 0:
            self._flush(0)
        for w in self.works:
            w.wait()
        proj = self.rows[:, self.blk:]
        if dtype != torch.float32:
            proj = proj.to(dtype)
        return [(self.w_edge(k), self.w_node(k), proj[k, : 3 * D * D].view(3 * D, D), proj[k, 3 * D * D:])
                for k in range(self.K)]


class MGNStackFn(torch.autograd.Function):
    """apply(cfg, plan, x, e_csr, *flat) with flat = (w_edge, w_node, w_proj, b_proj) per step."""

    @staticmethod
    def forward(ctx, cfg: StackConfig, plan: ops.GraphPlan, x: torch.Tensor, e: torch.Tensor, *flat):
        ops._require_cuda(x, e)
        if x.dtype != e.dtype:
            raise RuntimeError(f"node latents are {x.dtype} but edge latents are {e.dtype}")
        ops.dtype_code(x)
        if x.size(1) != D or e.size(1) != D:
            raise RuntimeError(f"the fused sm_100a path supports latent width {D} only (got {x.size(1)}, {e.size(1)})")
        if x.size(0) != plan.N or e.size(0) != plan.E:
            raise RuntimeError("latent shapes do not match the graph plan")
        K = len(flat) // 4
        x = x.contiguous()
        e = e.contiguous()
        path_e = ops.choose_path(x.dtype, cfg.act_edge, cfg.L_edge)
        path_n = ops.choose_path(x.dtype, cfg.act_node, cfg.L_node)
        scale = plan.inv_deg if cfg.mean else None
        paths_bwd = (ops.choose_path(x.dtype, cfg.act_edge, cfg.L_edge, backward=True),
                     ops.choose_path(x.dtype, cfg.act_node, cfg.L_node, backward=True))
        keep_h0 = ops.keeps_h0(path_e, path_n, *paths_bwd)
        keep_all = ops.keeps_hidden(keep_h0, cfg.L_edge, cfg.L_node, plan.E, plan.N, K, x.device)
        saved, preps = [], []
        for k in range(K):
            w_edge, w_node, w_proj, b_proj = flat[4 * k: 4 * k + 4]
            pe = ops.PreparedBlock(w_edge.detach(), cfg.L_edge, path_e, cfg.act_edge, cfg.use_ln)
            pn = ops.PreparedBlock(w_node.detach(), cfg.L_node, path_n, cfg.act_node, cfg.use_ln)
            preps.append((pe, pn))
            if ops.own_wgrad(x.dtype):   # P = x [W_s; W_d; W_nx]^T + b on the warp-specialised row-GEMM kernel
                P = ops.row_gemm([x], w_proj.detach(), w_mn=False, nb=3, bias=b_proj.detach())
            else:
                P = torch.addmm(b_proj.detach(), x, w_proj.detach().t())
            h0e = torch.empty_like(e) if keep_h0 else None
            h0n = torch.empty_like(x) if keep_h0 else None
            hhe = (torch.empty_like(e), torch.empty_like(e)) if keep_all else None     # H_1, H_2 of the edge block
            hhn = (torch.empty_like(x), torch.empty_like(x)) if keep_all else None
            e_new, agg = ops.block_fwd(pe, e, e, P, plan.src, plan.dst, 0, D, rowptr=plan.rowptr, want_agg=True,
                                       kind="edge_fwd", h0_out=h0e, hidden_out=hhe)
            # tcgen05 path: the node kernel also stores the rows its first GEMM consumed, round(agg * scale) in the
            # latent dtype -- what the backward's weight-gradient GEMM needs -- and the fp32 aggregate is not kept
            agg_lat = torch.empty_like(x) if (keep_h0 and x.dtype != torch.float32) else None
            x_new, _ = ops.block_fwd(pn, agg, x, P, None, None, 2 * D, 0, main_scale=scale, kind="node_fwd",
                                     h0_out=h0n, main_lat_out=agg_lat, hidden_out=hhn)
            # tcgen05 path: the first hidden activation of both blocks is kept (the backward then skips one gather,
            # one GEMM and one epilogue per tile and never reads P); CUDA-core path: P is kept and layer 0 recomputed
            saved += [x, e, agg_lat if agg_lat is not None else agg, h0e, h0n] if keep_h0 else [x, e, agg, P, P]
            if keep_all:    # keep-all policy: the backward kernels read H_1, H_2 instead of recomputing them
                saved += [hhe[0], hhe[1], hhn[0], hhn[1]]
            x, e = x_new, e_new
        ctx.cfg, ctx.plan, ctx.K = cfg, plan, K
        ctx.set_materialize_grads(False)
        ctx.paths, ctx.keep_h0, ctx.keep_all = paths_bwd, keep_h0, keep_all
        # the weight images of the forward serve the backward too when both run on the same kernel family
        ctx.preps = preps if (path_e, path_n) == paths_bwd else None
        ctx.save_for_backward(*saved, *flat)
        return x, e

    @staticmethod
    def backward(ctx, G_x, G_e):
        cfg, plan, K = ctx.cfg, ctx.plan, ctx.K
        path_e, path_n = ctx.paths
        saved = ctx.saved_tensors
        S = 9 if ctx.keep_all else 5                 # saved tensors per step
        acts, flat = saved[: S * K], saved[S * K:]
        dt = acts[0].dtype
        G_x = torch.zeros_like(acts[0]) if G_x is None else G_x.contiguous().to(dt)
        # G_e is updated in place layer by layer; the edge output is usually unused (no gradient materialised)
        G_e = torch.zeros_like(acts[1]) if G_e is None else G_e.contiguous().to(dt).clone()
        scale = plan.inv_deg if cfg.mean else None
        sink = GradSink(K, cfg.L_edge, cfg.L_node, acts[0].device)
        for k in reversed(range(K)):
            x, e, agg, a1, a2 = acts[S * k: S * k + 5]
            hhe, hhn = (acts[S * k + 5: S * k + 7], acts[S * k + 7: S * k + 9]) if ctx.keep_all else (None, None)
            P, h0e, h0n = (None, a1, a2) if ctx.keep_h0 else (a1, None, None)
            w_edge, w_node, w_proj, b_proj = flat[4 * k: 4 * k + 4]
            if ctx.preps is not None:
                pe, pn = ctx.preps[k]
            else:
                pe = ops.PreparedBlock(w_edge, cfg.L_edge, path_e, cfg.act_edge, cfg.use_ln)
                pn = ops.PreparedBlock(w_node, cfg.L_node, path_n, cfg.act_node, cfg.use_ln)
            # node block: g_agg, gradient of the node pre-activation, MLP weight grads
            lat = ctx.keep_h0 and agg.dtype != torch.float32      # agg = latent-dtype copy of the scaled aggregate
            g_agg, g_h0n, g_wn = ops.block_bwd(pn, agg, P, None, None, 2 * D, 0, G_x, main_scale=scale,
                                               kind="node_bwd", h0=h0n, n_nodes=plan.N, g_w_out=sink.w_node(k),
                                               main_is_lat_copy=lat, hidden=hhn)
            own = ops.own_wgrad(dt)                  # warp-specialised tcgen05 / TMA row reductions (csrc/wgrad.cu)
            agg_rows = agg if lat else (agg if scale is None else agg * scale[:, None]).to(dt)
            if own:
                ops.wgrad(g_h0n, agg_rows, g_wn[: D * D].view(D, D))
            else:
                ops.wgrad_into(g_wn, g_h0n, agg_rows)
            # edge block: total gradient of e' = G_e + g_agg[receiver]
            G_e, g_h0e, g_we = ops.block_bwd(pe, e, P, plan.src, plan.dst, 0, D, G_e, g_agg=g_agg,
                                             has_resid_grad=True, g_main_out=G_e, kind="edge_bwd", h0=h0e,
                                             n_nodes=plan.N, rowptr=plan.rowptr, g_w_out=sink.w_edge(k), hidden=hhe)
            # gradients of the gathered projections: segmented sums by sender and by receiver
            # (both land in one [N, 2D] matrix, so the products with W_s | W_d are single K = 2D GEMMs)
            g_psd = torch.empty((plan.N, 2 * D), dtype=dt, device=x.device)
            ops.segment_reduce(g_h0e, plan.sptr, plan.sperm, plan.N, out=g_psd[:, :D])
            if own:   # dW_e = g_h0e^T e and the receiver sums of g_h0e from one pass over the rows
                ops.wgrad(g_h0e, e, g_we[: D * D].view(D, D), seg=(plan.dst, plan.rowptr, plan.N, g_psd[:, D:]))
            else:
                ops.wgrad_into(g_we, g_h0e, e)
                ops.segment_reduce(g_h0e, plan.rowptr, None, plan.N, out=g_psd[:, D:])
            if own:   # g_x = [g_P_s | g_P_d | g_h0n] W + G_x: one K = 384 contraction, one rounding
                g_x = ops.row_gemm([g_psd[:, :D], g_psd[:, D:], g_h0n], w_proj, w_mn=True, add=G_x)
            else:
                g_x = torch.addmm(G_x, g_psd, w_proj[:2 * D])
                g_x.addmm_(g_h0n, w_proj[2 * D:])
            g_wproj = sink.w_proj(k)                 # fp32, written by the reductions
            if own:
                ops.wgrad(g_psd, x, g_wproj[:2 * D])
                ops.wgrad(g_h0n, x, g_wproj[2 * D:])
            else:
                torch.mm(g_psd.t(), x, out_dtype=torch.float32, out=g_wproj[:2 * D])
                torch.mm(g_h0n.t(), x, out_dtype=torch.float32, out=g_wproj[2 * D:])
            sink.step_done(k)
            G_x = g_x
        grads: List[Optional[torch.Tensor]] = []
        for per_step in sink.finish(flat[2].dtype):
            grads += list(per_step)
        return (None, None, G_x, G_e, *grads)


class _SegmentSumFn(torch.autograd.Function):
    """agg[n] = sum of rows [rowptr[n], rowptr[n+1]) (fp32, CSR order, deterministic); backward = row broadcast."""

    @staticmethod
    def forward(ctx, e, plan):
        ctx.plan = plan
        return ops.segment_reduce(e.contiguous(), plan.rowptr, None, plan.N, out_dtype=torch.float32)

    @staticmethod
    def backward(ctx, g):
        return ops.gather_rows(g.contiguous(), ctx.plan.dst), None


def _unpack_block(w: torch.Tensor, L: int):
    """(W_main, [(W_l, b_l)...], W_out, b_out, gamma, beta) views of a packed block vector (pack_block layout)."""
    mats = [w[i * D * D: (i + 1) * D * D].view(D, D) for i in range(L + 2)]
    v = w[(L + 2) * D * D:]
    vecs = [v[i * D: (i + 1) * D] for i in range(L + 3)]
    return mats[0], list(zip(mats[1: L + 1], vecs[:L])), mats[L + 1], vecs[L], vecs[L + 1], vecs[L + 2]


def eager_stack(cfg: StackConfig, plan: ops.GraphPlan, x: torch.Tensor, e: torch.Tensor, steps: Sequence[StepWeights]):
    """The same K processor steps as MGNStackFn as a chain of library ops ON THE GPU with torch autograd, for
    activations the fused kernels do not cover: they take an activation's derivative from its OUTPUT (relu, tanh,
    sigmoid, elu, leaky_relu), which SiLU / GELU / ... do not allow.  The reference accepts any torch.nn.functional
    name (mlp.py:37), so those run here -- announced and counted (ops.Fallbacks), never silently.  Receiver sums stay
    the deterministic segmented reduction."""
    import torch.nn.functional as F
    ops.Fallbacks.note("eager_activation", f"activation '{cfg.act_edge}'/'{cfg.act_node}' has no fused block kernel")
    fe, fn = getattr(F, cfg.act_edge), getattr(F, cfg.act_node)
    dt = x.dtype
    src, dst = plan.src.long(), plan.dst.long()

    def tail(h, act, hidden, w_out, b_out, gamma, beta):
        for w, b in hidden:
            h = act(F.linear(h, w.to(dt), b.to(dt)))
        y = F.linear(h, w_out.to(dt), b_out.to(dt))
        return F.layer_norm(y, (D,), gamma.to(dt), beta.to(dt), 1e-5) if cfg.use_ln else y

    for s in steps:
        we, he, woe, boe, ge, be = _unpack_block(s.w_edge, cfg.L_edge)
        wn, hn, won, bon, gn, bn = _unpack_block(s.w_node, cfg.L_node)
        P = F.linear(x, s.w_proj, s.b_proj)                                   # [N, 3D]
        h = fe(F.linear(e, we.to(dt)) + P[:, :D][src] + P[:, D:2 * D][dst])
        e = e + tail(h, fe, he, woe, boe, ge, be)
        agg = _SegmentSumFn.apply(e, plan)
        if cfg.mean:
            agg = agg * plan.inv_deg[:, None]
        h = fn(F.linear(agg.to(dt), wn.to(dt)) + P[:, 2 * D:])
        x = x + tail(h, fn, hn, won, bon, gn, bn)
    return x, e


def run_stack(cfg: StackConfig, plan: ops.GraphPlan, x: torch.Tensor, e_csr: torch.Tensor,
              steps: Sequence[StepWeights]):
    if cfg.act_edge not in ops._l.ACT_CODES or cfg.act_node not in ops._l.ACT_CODES:
        ops._require_cuda(x, e_csr)
        return eager_stack(cfg, plan, x, e_csr, steps)
    flat = []
    for s in steps:
        flat += [s.w_edge, s.w_node, s.w_proj, s.b_proj]
    return MGNStackFn.apply(cfg, plan, x, e_csr, *flat)


class PermuteRowsFn(torch.autograd.Function):
    """out[i] = inp[idx[i]] for a permutation idx (int32); backward gathers with the inverse."""

    @staticmethod
    def forward(ctx, inp, idx, inv_idx):
        ctx.inv_idx = inv_idx
        return ops.gather_rows(inp, idx)

    @staticmethod
    def backward(ctx, g):
        return ops.gather_rows(g.contiguous(), ctx.inv_idx), None, None


def permute_rows(inp: torch.Tensor, idx: torch.Tensor, inv_idx: torch.Tensor) -> torch.Tensor:
    return PermuteRowsFn.apply(inp, idx, inv_idx)


class SingleBlockFn(torch.autograd.Function):
    """One fused block without residual: the standalone EdgeBlock / EdgeBlockSum / NodeBlock forward
    (mgnLayer.py:32-49, :93-105, :134-153).  `e` is in receiver-CSR order."""

    @staticmethod
    def forward(ctx, mode: str, L: int, act: str, use_ln: bool, mean: bool, plan: ops.GraphPlan, e, x, w, w_proj,
                b_proj):
        ops._require_cuda(e, x)
        if x.dtype != e.dtype:
            raise RuntimeError(f"node latents are {x.dtype} but edge latents are {e.dtype}")
        if x.size(1) != D or e.size(1) != D:
            raise RuntimeError(f"the fused sm_100a path supports latent width {D} only")
        e, x = e.contiguous(), x.contiguous()
        path = ops.choose_path(x.dtype, act, L)
        prep = ops.PreparedBlock(w.detach(), L, path, act, use_ln)
        path = ops.choose_path(x.dtype, act, L, backward=True)
        P = torch.addmm(b_proj.detach(), x, w_proj.detach().t())
        scale = plan.inv_deg if (mean and mode == "node") else None
        if mode == "edge":
            agg = None
            out, _ = ops.block_fwd(prep, e, None, P, plan.src, plan.dst, 0, D)
        else:
            agg = ops.segment_reduce(e, plan.rowptr, None, plan.N, out_dtype=torch.float32)
            out, _ = ops.block_fwd(prep, agg, None, P, None, None, 0, 0, main_scale=scale)
        ctx.meta = (mode, L, act, use_ln, path, plan, scale)
        ctx.save_for_backward(e, x, agg if agg is not None else e.new_empty(0), w, w_proj, b_proj)
        return out

    @staticmethod
    def backward(ctx, g):
        mode, L, act, use_ln, path, plan, scale = ctx.meta
        e, x, agg, w, w_proj, b_proj = ctx.saved_tensors
        dt = x.dtype
        g = g.contiguous().to(dt)
        prep = ops.PreparedBlock(w, L, path, act, use_ln)
        P = torch.addmm(b_proj, x, w_proj.t())
        if mode == "edge":
            g_e, g_h0, g_w = ops.block_bwd(prep, e, P, plan.src, plan.dst, 0, D, g)
            ops.wgrad_into(g_w, g_h0, e)
            g_ps = ops.segment_reduce(g_h0, plan.sptr, plan.sperm, plan.N)
            g_pd = ops.segment_reduce(g_h0, plan.rowptr, None, plan.N)
            g_x = torch.addmm(g_ps @ w_proj[:D], g_pd, w_proj[D:])
            g_wproj = torch.cat([g_ps.t() @ x, g_pd.t() @ x], dim=0)
            g_b = torch.cat([g_w[-D:], g_w[-D:]]).to(b_proj.dtype)
        else:
            g_agg, g_h0, g_w = ops.block_bwd(prep, agg, P, None, None, 0, 0, g, main_scale=scale)
            agg_eff = agg if scale is None else agg * scale[:, None]
            ops.wgrad_into(g_w, g_h0, agg_eff.to(dt))
            g_e = ops.gather_rows(g_agg.to(dt), plan.dst)
            g_x = g_h0 @ w_proj
            g_wproj = g_h0.t() @ x
            g_b = g_w[-D:].to(b_proj.dtype)
        return (None,) * 6 + (g_e, g_x, g_w, g_wproj.to(w_proj.dtype), g_b)


def _plan_and_csr(edge_attr, node_attr, edge_index):
    ops._require_cuda(node_attr, edge_attr, edge_index)
    plan = ops.PLAN_CACHE.get(edge_index, node_attr.size(0))
    return plan, permute_rows(edge_attr, plan.perm, plan.inv_perm)


def edge_block_apply(parts: dict, edge_attr, node_attr, edge_index):
    """Standalone edge block: returns the edge update in the caller's edge order (no residual)."""
    plan, e_csr = _plan_and_csr(edge_attr, node_attr, edge_index)
    dt = node_attr.dtype
    w = pack_block(parts["w_e"], parts["hidden"], parts["w_out"], parts["b_out"], parts["gamma"], parts["beta"])
    w_proj = torch.cat([parts["w_s"], parts["w_d"]], dim=0).to(dt)
    b_proj = torch.cat([torch.zeros_like(parts["b0"]), parts["b0"]]).to(dt)
    u = SingleBlockFn.apply("edge", len(parts["hidden"]), parts["act"], parts["use_ln"], False, plan, e_csr, node_attr,
                            w, w_proj, b_proj)
    return permute_rows(u, plan.inv_perm, plan.perm)


def node_block_apply(parts: dict, node_attr, edge_attr, edge_index, mean: bool):
    """Standalone node block: aggregate incoming edge rows, return the node update (no residual)."""
    plan, e_csr = _plan_and_csr(edge_attr, node_attr, edge_index)
    dt = node_attr.dtype
    w = pack_block(parts["w_a"], parts["hidden"], parts["w_out"], parts["b_out"], parts["gamma"], parts["beta"])
    return SingleBlockFn.apply("node", len(parts["hidden"]), parts["act"], parts["use_ln"], mean, plan, e_csr, node_attr,
                               w, parts["w_x"].to(dt), parts["b0"].to(dt))


# ------------------------------------------------------------------------------------------------
# standalone MLP (encoders): everything after the first Linear on the fused block kernel
# ------------------------------------------------------------------------------------------------
_ZERO_IDX: dict = {}
_EYE: dict = {}


def _zero_idx(rows: int, device) -> torch.Tensor:
    """int32 zeros [rows]: every row gathers row 0 of the (single-row) pre-projection."""
    key = (device.index, rows)
    t = _ZERO_IDX.get(key)
    if t is None:
        if len(_ZERO_IDX) > 8:
            _ZERO_IDX.clear()
        t = _ZERO_IDX[key] = torch.zeros(rows, dtype=torch.int32, device=device)
    return t


class DenseTailFn(torch.autograd.Function):
    """out = LN(W_out act(.. act(W_1 act(z) + b_1) ..) + b_out) for z = the first Linear's output, bias included
    (reference models/mlp.py:40-51 after its first `layer(x)`).  The block kernel's first GEMM runs with
    W_main = I and a zero pre-projection row, so h_0 = act(z); `w` is the packed vector of pack_block().  On the
    tcgen05 path h_0 is kept, so the backward is the TMA-fed kernel (no recompute of layer 0)."""

    @staticmethod
    def forward(ctx, L: int, act: str, use_ln: bool, z: torch.Tensor, w: torch.Tensor):
        ops._require_cuda(z, w)
        z = z.contiguous()
        P = torch.zeros((1, D), dtype=z.dtype, device=z.device)
        idx0 = _zero_idx(z.size(0), z.device)
        path = ops.choose_path(z.dtype, act, L)
        path_b = ops.choose_path(z.dtype, act, L, backward=True)
        prep = ops.PreparedBlock(w.detach(), L, path, act, use_ln)
        h0 = torch.empty_like(z) if (ops.keeps_h0(path, path_b) and 1 <= L <= 2 and z.size(0) > 0) else None
        out, _ = ops.block_fwd(prep, z, None, P, idx0, None, 0, 0, kind="dense_fwd", h0_out=h0)
        ctx.meta = (L, act, use_ln, path_b)
        ctx.has_h0 = h0 is not None
        ctx.save_for_backward(z if h0 is None else h0, w, P, idx0)
        return out

    @staticmethod
    def backward(ctx, g):
        L, act, use_ln, path = ctx.meta
        z, w, P, idx0 = ctx.saved_tensors          # z: the rows themselves, or the kept h_0 (z is then not needed)
        prep = ops.PreparedBlock(w, L, path, act, use_ln)
        # W_main = I: the gradient of the first pre-activation IS the gradient of z (g_main = g_h0 . I is ignored)
        _, g_h0, g_w = ops.block_bwd(prep, z, P, idx0, None, 0, 0, g.contiguous().to(z.dtype), kind="dense_bwd",
                                     h0=z if ctx.has_h0 else None)
        return None, None, None, g_h0, g_w


def dense_tail(L: int, act: str, use_ln: bool, z: torch.Tensor, hidden: Sequence, w_out, b_out, gamma, beta):
    key = (z.device.index,)
    eye = _EYE.get(key)
    if eye is None:
        eye = _EYE[key] = torch.eye(D, dtype=torch.float32, device=z.device)
    w = pack_block(eye, hidden, w_out, b_out, gamma, beta)
    return DenseTailFn.apply(L, act, use_ln, z, w)


class DenseMLPFn(torch.autograd.Function):
    """A whole MLP whose input is D wide (the decoder, mgn.py:130): out = [LN](W_out act(.. act(W_0 x + b_0) ..) + b_out).
    W_0 is the block's real first GEMM, b_0 its single pre-projection row; a narrower last Linear is zero-padded to
    D output columns by the caller (the padded columns are zeros and carry zero gradient).  h_0 is kept on the
    tcgen05 path (TMA-fed backward)."""

    @staticmethod
    def forward(ctx, L: int, act: str, use_ln: bool, x: torch.Tensor, w: torch.Tensor, b0: torch.Tensor):
        ops._require_cuda(x, w, b0)
        x = x.contiguous()
        P = b0.detach().to(x.dtype).reshape(1, D).contiguous()
        idx0 = _zero_idx(x.size(0), x.device)
        path = ops.choose_path(x.dtype, act, L)
        path_b = ops.choose_path(x.dtype, act, L, backward=True)
        prep = ops.PreparedBlock(w.detach(), L, path, act, use_ln)
        h0 = torch.empty_like(x) if (ops.keeps_h0(path, path_b) and 1 <= L <= 2 and x.size(0) > 0) else None
        out, _ = ops.block_fwd(prep, x, None, P, idx0, None, 0, 0, kind="dense_fwd", h0_out=h0)
        ctx.meta = (L, act, use_ln, path_b, b0.dtype)
        ctx.save_for_backward(x, w, P, idx0, h0 if h0 is not None else x.new_empty(0))
        return out

    @staticmethod
    def backward(ctx, g):
        L, act, use_ln, path, b0_dtype = ctx.meta
        x, w, P, idx0, h0 = ctx.saved_tensors
        prep = ops.PreparedBlock(w, L, path, act, use_ln)
        g_x, g_h0, g_w = ops.block_bwd(prep, x, P, idx0, None, 0, 0, g.contiguous().to(x.dtype), kind="dense_bwd",
                                       h0=h0 if h0.numel() else None)
        if ops.own_wgrad(x.dtype) and x.size(0) > 0:     # dW_0 = g_h0^T x, like the processor's first-layer gradients
            ops.wgrad(g_h0, x, g_w[: D * D].view(D, D))
        else:
            ops.wgrad_into(g_w, g_h0, x)
        return None, None, None, g_x, g_w, g_w[-D:].to(b0_dtype)   # last slot: column sums of g_h0 = d b_0


def dense_mlp(L: int, act: str, use_ln: bool, x: torch.Tensor, w0, b0, hidden: Sequence, w_out, b_out, gamma, beta):
    """x [rows, D] through a full MLP on the fused block kernel; returns [rows, D] (columns >= w_out.size(0) are 0)."""
    out_dim = w_out.size(0)
    if out_dim < D:
        pad = D - out_dim
        w_out = torch.cat([w_out, w_out.new_zeros((pad, D))], dim=0)
        b_out = torch.cat([b_out, b_out.new_zeros(pad)])
        gamma = torch.cat([gamma, gamma.new_ones(pad)])
        beta = torch.cat([beta, beta.new_zeros(pad)])
    w = pack_block(w0, hidden, w_out, b_out, gamma, beta)
    return DenseMLPFn.apply(L, act, use_ln, x, w, b0)
