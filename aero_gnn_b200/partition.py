"""One large mesh across the GPUs of a box: contiguous receiver-node blocks + one halo exchange per step.

The reference has no distributed code; this is new.  Rank r owns nodes [lo, hi) and every edge
whose receiver it owns, so edge latents, the edge block and the receiver sums stay local and deterministic.  A
processor step reads only 1-hop sender latents (mgnLayer.py:40-41), so the only data crossing ranks is the latent
row of each remote sender ("halo"), once per step forward and the gradient of those rows once per step backward.
Weights are replicated; their gradients are summed by all-reduces of flat fp32 buckets (5 steps each) that travel
under the backward of the earlier steps.

Local numbering.  Own nodes first, halo nodes (ascending global id) after them.  Inside the own block the
INTERIOR receivers -- own nodes all of whose incoming edges have an owned sender -- come first and the BOUNDARY
receivers (at least one remote sender) last, each group in ascending global id.  Edges are kept in receiver-CSR
order of the local numbering, so the edges of interior receivers are the row range [0, E_int) and need nothing from
other ranks: the edge kernel runs over them while the halo rows are still travelling, and over [E_int, E_loc) after
they arrived (north star: "halo ... exchanged by NCCL over NVLink once per message-passing step and overlapped with
interior edges").  Because owners hold contiguous id ranges, the halo rows coming from one peer are a contiguous
slice of the halo block.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .processor import D, GradSink, StackConfig


def block_bounds(n_nodes: int, world: int, rank: int):
    blk = -(-n_nodes // world)
    lo = min(rank * blk, n_nodes)
    return lo, min(lo + blk, n_nodes)


@dataclass
class HaloPlan:
    """Pure index plan (numpy / CPU): which rows to send to / receive from each peer."""

    rank: int
    world: int
    lo: int
    hi: int
    edge_ids: np.ndarray            # global caller edge ids owned by this rank (ascending)
    local_edge_index: np.ndarray    # [2, E_loc] int64: local sender id, local receiver id
    halo_global: np.ndarray         # [n_halo] ascending global ids of remote senders
    recv_counts: List[int]          # rows received from each peer (contiguous slices of the halo block)
    send_idx: List[np.ndarray]      # local row ids of the rows each peer needs, in ascending global id
    own_order: np.ndarray           # [n_own] global id - lo stored at local row i (interior receivers first)
    n_interior: int                 # local rows [0, n_interior) receive from owned senders only

    @property
    def n_own(self) -> int:
        return self.hi - self.lo

    @property
    def n_halo(self) -> int:
        return int(self.halo_global.shape[0])

    @property
    def n_local(self) -> int:
        return self.n_own + self.n_halo


def build_halo_plan(edge_index: np.ndarray, n_nodes: int, rank: int, world: int, reorder: bool = True) -> HaloPlan:
    """Every rank holds the full connectivity (16E bytes), so the plan needs no communication.  `reorder=False`
    keeps the own rows in global order (n_interior = 0: no interior / boundary split)."""
    src, dst = edge_index[0], edge_index[1]
    lo, hi = block_bounds(n_nodes, world, rank)
    n_own = hi - lo
    mine = np.flatnonzero((dst >= lo) & (dst < hi))
    s, d = src[mine], dst[mine] - lo
    remote = (s < lo) | (s >= hi)
    halo = np.unique(s[remote])
    if reorder:
        boundary = np.zeros(n_own, dtype=bool)
        boundary[d[remote]] = True
        own_order = np.concatenate([np.flatnonzero(~boundary), np.flatnonzero(boundary)]).astype(np.int64)
        n_interior = int(n_own - boundary.sum())
    else:
        own_order, n_interior = np.arange(n_own, dtype=np.int64), 0
    new_of_old = np.empty(n_own, dtype=np.int64)
    new_of_old[own_order] = np.arange(n_own)
    loc = np.where(remote, n_own + np.searchsorted(halo, s), new_of_old[np.where(remote, 0, s - lo)] if n_own else 0)
    recv_counts, send_idx = [], []
    for p in range(world):
        plo, phi = block_bounds(n_nodes, world, p)
        recv_counts.append(int(np.count_nonzero((halo >= plo) & (halo < phi))) if p != rank else 0)
        if p == rank:
            send_idx.append(np.empty(0, dtype=np.int64))
            continue
        theirs = (dst >= plo) & (dst < phi) & (src >= lo) & (src < hi)      # p's edges whose sender I own
        send_idx.append(new_of_old[np.unique(src[theirs]) - lo])
    led = np.stack([loc, new_of_old[d] if n_own else d]).astype(np.int64)
    return HaloPlan(rank, world, lo, hi, mine, led, halo, recv_counts, send_idx, own_order, n_interior)


class HaloExchanger:
    """Point-to-point halo exchange over torch.distributed (NCCL on GPUs, gloo in the CPU tests).

    One packed buffer per direction: the rows every peer needs are gathered with ONE launch into a send buffer whose
    per-peer slices are posted, receives land straight in the halo block of the extended row matrix; in the
    backward the returned rows arrive in one buffer and are added to their owners by one segmented reduction in
    ascending peer order (deterministic) instead of one index_add_ per peer."""

    def __init__(self, plan: HaloPlan, device, group=None):
        self.plan, self.group = plan, group
        self.device = torch.device(device)
        self.recv_off = np.concatenate([[0], np.cumsum(plan.recv_counts)]).astype(int)
        counts = [int(ix.shape[0]) for ix in plan.send_idx]
        self.send_off = np.concatenate([[0], np.cumsum(counts)]).astype(int)
        cat = np.concatenate(plan.send_idx) if plan.send_idx else np.empty(0, dtype=np.int64)
        self.n_send = int(cat.shape[0])
        self.send_rows = torch.from_numpy(cat.astype(np.int64)).to(self.device)             # local row per send slot
        self.send_rows32 = self.send_rows.to(torch.int32)
        # owners' side of the backward: unique own rows, and for each the send slots that return a gradient for it
        uniq, inv = np.unique(cat, return_inverse=True)
        order = np.argsort(inv, kind="stable")                                               # slots grouped by row, peer order kept
        ptr = np.concatenate([[0], np.cumsum(np.bincount(inv, minlength=uniq.shape[0]))]) if uniq.size else np.zeros(1)
        self.ret_rows = torch.from_numpy(uniq.astype(np.int64)).to(self.device)
        self.ret_ptr = torch.from_numpy(ptr.astype(np.int32)).to(self.device)
        self.ret_list = torch.from_numpy(order.astype(np.int32)).to(self.device)

    def _post(self, sends: Sequence[Optional[torch.Tensor]], recvs: Sequence[Optional[torch.Tensor]]):
        """Post all receives and sends; returns the outstanding requests (the buffers must outlive them)."""
        p2p = []
        for p in range(self.plan.world):
            if recvs[p] is not None and recvs[p].numel():
                p2p.append(dist.P2POp(dist.irecv, recvs[p], p, self.group))
        for p in range(self.plan.world):
            if sends[p] is not None and sends[p].numel():
                p2p.append(dist.P2POp(dist.isend, sends[p], p, self.group))
        return dist.batch_isend_irecv(p2p) if p2p else []

    def _slices(self, buf: torch.Tensor, off) -> List[Optional[torch.Tensor]]:
        return [buf[off[p]: off[p + 1]] if off[p + 1] > off[p] else None for p in range(self.plan.world)]

    # ---- split-phase API: post, do independent work on the compute stream, then finish ----------------------
    def forward_start(self, x_own: torch.Tensor, out: Optional[torch.Tensor] = None):
        """Post the exchange of the remote senders' rows into `out` ([n_halo, width], halo order; allocated when
        None).  `x_own`: the own rows in local order (any tensor whose first n_own rows are them).  Returns a token
        for forward_finish(); `out` is valid only after it."""
        pl = self.plan
        halo = out if out is not None else x_own.new_empty((pl.n_halo, x_own.size(1)))
        if self.n_send == 0:
            send = x_own.new_empty((0, x_own.size(1)))
        elif x_own.is_cuda and x_own.dtype in (torch.float32, torch.bfloat16):
            send = ops.gather_rows(x_own, self.send_rows32)          # one launch for every peer's rows
        else:
            send = x_own[self.send_rows]
        return halo, send, self._post(self._slices(send, self.send_off), self._slices(halo, self.recv_off))

    @staticmethod
    def forward_finish(token) -> torch.Tensor:
        halo, _send, reqs = token
        for req in reqs:
            req.wait()
        return halo

    def backward_start(self, g_halo: torch.Tensor, like: torch.Tensor):
        """Post the return of the halo-row gradients to their owners (`like`: a tensor of the owners' dtype/device).
        `g_halo`: [n_halo, width] (a row slice of a contiguous matrix is sent without a copy)."""
        g_halo = g_halo.contiguous()
        recv = like.new_empty((self.n_send, like.size(1)))
        return g_halo, recv, self._post(self._slices(g_halo, self.recv_off), self._slices(recv, self.send_off))

    def backward_finish(self, token, g_own: torch.Tensor) -> None:
        """Owners add the returned rows: per own row the contributions are summed in ascending peer order (fp32
        accumulation, fixed order -> deterministic), then added to the row."""
        _g_halo, recv, reqs = token
        for req in reqs:
            req.wait()
        if self.n_send == 0:
            return
        n_ret = int(self.ret_rows.numel())
        if recv.is_cuda and recv.dtype in (torch.float32, torch.bfloat16):
            tot = ops.segment_reduce(recv, self.ret_ptr, self.ret_list, n_ret)
        else:   # host path of the gloo tests: same order, plain torch
            counts = (self.ret_ptr[1:] - self.ret_ptr[:-1]).long()
            seg = torch.repeat_interleave(torch.arange(n_ret, device=recv.device), counts)
            tot = torch.zeros((n_ret, recv.size(1)), dtype=recv.dtype, device=recv.device)
            tot.index_add_(0, seg, recv[self.ret_list.long()])
        g_own.index_add_(0, self.ret_rows, tot)     # distinct rows: no two updates meet

    # ---- blocking forms ------------------------------------------------------------------------------------
    def forward(self, x_own: torch.Tensor) -> torch.Tensor:
        """Rows of the remote senders, [n_halo, width], in halo order."""
        return self.forward_finish(self.forward_start(x_own))

    def backward(self, g_halo: torch.Tensor, g_own: torch.Tensor) -> None:
        """Send halo-row gradients back to their owners and accumulate them into g_own."""
        self.backward_finish(self.backward_start(g_halo, g_own), g_own)


class PartitionedStackFn(torch.autograd.Function):
    """MGN processor stack on one receiver block; apply(cfg, part, x_own, e_csr, *flat) like processor.MGNStackFn.
    `x_own` is in LOCAL own-row order (PartitionedProcessor.run permutes)."""

    @staticmethod
    def forward(ctx, cfg: StackConfig, part: "PartitionedProcessor", x, e, *flat):
        plan, ex = part.plan, part.exchanger
        n_own, E_int = part.n_own, part.E_int
        K = len(flat) // 4
        x, e = x.contiguous(), e.contiguous()
        dt, dev = x.dtype, x.device
        path_e = ops.choose_path(dt, cfg.act_edge, cfg.L_edge)
        path_n = ops.choose_path(dt, cfg.act_node, cfg.L_node)
        scale = part.inv_deg_own if cfg.mean else None
        paths_bwd = (ops.choose_path(dt, cfg.act_edge, cfg.L_edge, backward=True),
                     ops.choose_path(dt, cfg.act_node, cfg.L_node, backward=True))
        keep_h0 = ops.keeps_h0(path_e, path_n, *paths_bwd)
        keep_all = ops.keeps_hidden(keep_h0, cfg.L_edge, cfg.L_node, plan.E, n_own, K, x.device)
        split = 0 < E_int < plan.E           # interior edges first, boundary edges after the halo arrived
        saved, preps = [], []
        x_ext = x.new_empty((plan.N, D))
        x_ext[:n_own].copy_(x)
        x_new = x
        for k in range(K):
            w_edge, w_node, w_proj, b_proj = flat[4 * k: 4 * k + 4]
            pe = ops.PreparedBlock(w_edge.detach(), cfg.L_edge, path_e, cfg.act_edge, cfg.use_ln)
            pn = ops.PreparedBlock(w_node.detach(), cfg.L_node, path_n, cfg.act_node, cfg.use_ln)
            preps.append((pe, pn))
            x_cur = x_ext[:n_own]
            # the halo rows travel while the own rows are pre-projected and the interior edges are processed
            tok = ex.forward_start(x_ext, out=x_ext[n_own:])
            P = x.new_empty((plan.N, w_proj.size(0)))
            wt, bp = w_proj.detach().t(), b_proj.detach()
            own_rows = ops.own_wgrad(x.dtype)          # row GEMMs on csrc/rowgemm.cu instead of the library
            if own_rows:
                ops.row_gemm([x_cur], w_proj.detach(), w_mn=False, nb=3, bias=bp, out=P[:n_own])
            else:
                torch.addmm(bp, x_cur, wt, out=P[:n_own])
            h0e = torch.empty_like(e) if keep_h0 else None
            h0n = torch.empty_like(x) if keep_h0 else None
            hhe = (torch.empty_like(e), torch.empty_like(e)) if keep_all else None
            hhn = (torch.empty_like(x), torch.empty_like(x)) if keep_all else None
            e_new = torch.empty_like(e)
            agg_full = torch.empty((plan.N, D), dtype=torch.float32, device=dev)
            if split:
                ops.block_fwd(pe, e, e, P, plan.src, plan.dst, 0, D, rowptr=plan.rowptr, kind="edge_fwd",
                              h0_out=h0e, out=e_new, agg_out=agg_full, rows=(0, E_int), hidden_out=hhe)
            ex.forward_finish(tok)
            if plan.N > n_own:
                if own_rows:
                    ops.row_gemm([x_ext[n_own:]], w_proj.detach(), w_mn=False, nb=3, bias=bp, out=P[n_own:])
                else:
                    torch.addmm(bp, x_ext[n_own:], wt, out=P[n_own:])
            if split:
                ops.block_fwd(pe, e, e, P, plan.src, plan.dst, 0, D, rowptr=part.rowptr_boundary, kind="edge_fwd_b",
                              h0_out=h0e, out=e_new, agg_out=agg_full, agg_clear=False, rows=(E_int, plan.E),
                              hidden_out=hhe)
            else:
                ops.block_fwd(pe, e, e, P, plan.src, plan.dst, 0, D, rowptr=plan.rowptr, kind="edge_fwd",
                              h0_out=h0e, out=e_new, agg_out=agg_full, hidden_out=hhe)
            agg = agg_full[:n_own]
            # x' goes straight into the own rows of the next step's extended row matrix
            x_next = x.new_empty((plan.N, D)) if k + 1 < K else None
            out_rows = x_next[:n_own] if x_next is not None else None
            agg_lat = torch.empty_like(x) if (keep_h0 and dt != torch.float32) else None
            x_new, _ = ops.block_fwd(pn, agg, x_cur, P, None, None, 2 * D, 0, main_scale=scale, kind="node_fwd",
                                     h0_out=h0n, out=out_rows, main_lat_out=agg_lat, hidden_out=hhn)
            # x_ext and (h_0 of both blocks | P) are kept so the backward needs no second halo exchange; with kept h_0
            # the aggregate is kept as the latent-dtype copy the node kernel made of its staged rows
            saved += [x_ext, e, agg_lat if agg_lat is not None else agg, h0e, h0n] if keep_h0 else [x_ext, e, agg, P, P]
            if keep_all:
                saved += [hhe[0], hhe[1], hhn[0], hhn[1]]
            e = e_new
            if x_next is not None:
                x_ext = x_next
        ctx.cfg, ctx.part, ctx.K = cfg, part, K
        ctx.set_materialize_grads(False)
        ctx.paths, ctx.keep_h0, ctx.keep_all = paths_bwd, keep_h0, keep_all
        # the weight images of the forward serve the backward too when both run on the same kernel family
        ctx.preps = preps if (path_e, path_n) == paths_bwd else None
        ctx.save_for_backward(*saved, *flat)
        return x_new, e

    @staticmethod
    def backward(ctx, G_x, G_e):
        cfg, part, K = ctx.cfg, ctx.part, ctx.K
        plan, ex, n_own = part.plan, part.exchanger, part.n_own
        path_e, path_n = ctx.paths
        saved = ctx.saved_tensors
        S = 9 if ctx.keep_all else 5
        acts, flat = saved[: S * K], saved[S * K:]
        dt = acts[0].dtype
        G_x = torch.zeros_like(acts[0][:n_own]) if G_x is None else G_x.contiguous().to(dt)
        G_e = torch.zeros_like(acts[1]) if G_e is None else G_e.contiguous().to(dt).clone()
        scale = part.inv_deg_own if cfg.mean else None
        # parameter gradients: flat fp32 buckets, summed over the ranks while earlier steps are still in their backward
        reduce = (lambda t: dist.all_reduce(t, group=part.group, async_op=True)) if part.world > 1 else None
        sink = GradSink(K, cfg.L_edge, cfg.L_node, acts[0].device, reduce=reduce)
        for k in reversed(range(K)):
            x_ext, e, agg, a1, a2 = acts[S * k: S * k + 5]
            hhe, hhn = (acts[S * k + 5: S * k + 7], acts[S * k + 7: S * k + 9]) if ctx.keep_all else (None, None)
            P, h0e, h0n = (None, a1, a2) if ctx.keep_h0 else (a1, None, None)
            x = x_ext[:n_own]
            w_edge, w_node, w_proj, b_proj = flat[4 * k: 4 * k + 4]
            if ctx.preps is not None:
                pe, pn = ctx.preps[k]
            else:
                pe = ops.PreparedBlock(w_edge, cfg.L_edge, path_e, cfg.act_edge, cfg.use_ln)
                pn = ops.PreparedBlock(w_node, cfg.L_node, path_n, cfg.act_node, cfg.use_ln)
            lat = ctx.keep_h0 and agg.dtype != torch.float32
            g_agg, g_h0n, g_wn = ops.block_bwd(pn, agg, P, None, None, 2 * D, 0, G_x, main_scale=scale, kind="node_bwd",
                                               h0=h0n, n_nodes=plan.N, g_w_out=sink.w_node(k), main_is_lat_copy=lat,
                                               hidden=hhn)
            own = ops.own_wgrad(dt)                  # warp-specialised tcgen05 / TMA row reductions (csrc/wgrad.cu)
            agg_rows = agg if lat else (agg if scale is None else agg * scale[:, None]).to(dt)
            if own:
                ops.wgrad(g_h0n, agg_rows, g_wn[: D * D].view(D, D))
            else:
                ops.wgrad_into(g_wn, g_h0n, agg_rows)
            G_e, g_h0e, g_we = ops.block_bwd(pe, e, P, plan.src, plan.dst, 0, D, G_e, g_agg=g_agg, has_resid_grad=True,
                                             g_main_out=G_e, kind="edge_bwd", h0=h0e, n_nodes=plan.N,
                                             rowptr=plan.rowptr, g_w_out=sink.w_edge(k), hidden=hhe)
            g_psd = torch.empty((plan.N, 2 * D), dtype=dt, device=e.device)   # [g_P_s | g_P_d] over local rows
            ops.segment_reduce(g_h0e, plan.sptr, plan.sperm, plan.N, out=g_psd[:, :D])
            if own:
                ops.wgrad(g_h0e, e, g_we[: D * D].view(D, D), seg=(plan.dst, plan.rowptr, plan.N, g_psd[:, D:]))
            else:
                ops.wgrad_into(g_we, g_h0e, e)
                ops.segment_reduce(g_h0e, plan.rowptr, None, plan.N, out=g_psd[:, D:])
            if own:
                # halo rows first (only the gathered projections reach them), so their gradients travel under the
                # own rows' K = 384 contraction
                g_halo = ops.row_gemm([g_psd[n_own:, :D], g_psd[n_own:, D:]], w_proj[:2 * D], w_mn=True)
                tok = ex.backward_start(g_halo, g_halo)
                g_x = ops.row_gemm([g_psd[:n_own, :D], g_psd[:n_own, D:], g_h0n], w_proj, w_mn=True, add=G_x)
            else:
                g_ext = g_psd @ w_proj[:2 * D]                          # [n_local, D]
                tok = ex.backward_start(g_ext[n_own:], g_ext)           # halo-row gradients travel under the GEMMs below
                g_x = G_x + g_ext[:n_own]
                g_x.addmm_(g_h0n, w_proj[2 * D:])
            g_wproj = sink.w_proj(k)
            if own:
                ops.wgrad(g_psd, x_ext, g_wproj[:2 * D])
                ops.wgrad(g_h0n, x, g_wproj[2 * D:])
            else:
                torch.mm(g_psd.t(), x_ext, out_dtype=torch.float32, out=g_wproj[:2 * D])
                torch.mm(g_h0n.t(), x, out_dtype=torch.float32, out=g_wproj[2 * D:])
            ex.backward_finish(tok, g_x)
            sink.step_done(k)
            G_x = g_x
        grads = []
        for per_step in sink.finish(flat[2].dtype):
            grads += list(per_step)
        return (None, None, G_x, G_e, *grads)


class HaloExtendFn(torch.autograd.Function):
    """x_ext = [own rows (local order) | halo rows fetched from their owners]; the backward returns the halo rows'
    gradients to the owners and adds them there (deterministic order).  Lets any 1-hop operator that works on a whole
    graph -- WeightedEdgeConv of the BFS-bistride variant, SURVEY.md 8(e) row 3 -- run on a rank's local graph."""

    @staticmethod
    def forward(ctx, part: "PartitionedProcessor", x_own: torch.Tensor):
        ex, n_own = part.exchanger, part.n_own
        x_ext = x_own.new_empty((part.plan.N, x_own.size(1)))
        x_ext[:n_own].copy_(x_own)
        ex.forward_finish(ex.forward_start(x_ext, out=x_ext[n_own:]))
        ctx.part = part
        return x_ext

    @staticmethod
    def backward(ctx, g_ext):
        part = ctx.part
        n_own = part.n_own
        g_ext = g_ext.contiguous()
        g_own = g_ext[:n_own].clone()
        part.exchanger.backward(g_ext[n_own:], g_own)
        return None, g_own


class PartitionedProcessor:
    """Receiver-block partition of one mesh for this rank: halo plan, local graph plan, exchange, grad all-reduce."""

    def __init__(self, edge_index: torch.Tensor, n_nodes: int, rank: int, world: int, device, group=None,
                 overlap: bool = True):
        ei = edge_index.cpu().numpy()
        self.halo = build_halo_plan(ei, n_nodes, rank, world, reorder=overlap)
        self.rank, self.world, self.group = rank, world, group
        self.lo, self.hi, self.n_own = self.halo.lo, self.halo.hi, self.halo.n_own
        self.edge_ids_cpu = torch.from_numpy(self.halo.edge_ids)
        local_ei = torch.from_numpy(self.halo.local_edge_index).to(device)
        self.plan = ops.build_graph_plan(local_ei, self.halo.n_local)
        self.E_loc = self.plan.E
        self.exchanger = HaloExchanger(self.halo, device, group)
        # local own-row order <-> global order (own rows only)
        order = torch.from_numpy(self.halo.own_order).to(device)
        self.own_order = order.to(torch.int32)                       # local row i holds global own row own_order[i]
        inv = torch.empty_like(order)
        inv[order] = torch.arange(order.numel(), device=order.device)
        self.own_inv = inv.to(torch.int32)
        self.identity_order = bool((self.halo.own_order == np.arange(self.n_own)).all())
        # edges of interior receivers = CSR rows [0, E_int); receiver CSR relative to E_int for the second launch
        self.E_int = int(self.plan.rowptr[self.halo.n_interior].item())
        self.rowptr_boundary = (self.plan.rowptr - self.E_int).contiguous()
        self.inv_deg_own = self.plan.inv_deg[: self.n_own].contiguous()
        self._stack_params = set()      # ids of the parameters whose gradients the stack's backward already summed

    def csr_edge_ids(self) -> torch.Tensor:
        """Global caller edge id stored at each local CSR slot."""
        return self.edge_ids_cpu.to(self.plan.perm.device)[self.plan.perm.long()]

    def to_local(self, x_own_global_order: torch.Tensor) -> torch.Tensor:
        """Own rows from ascending-global-id order to the local (interior first) order."""
        from .processor import permute_rows
        if self.identity_order:
            return x_own_global_order
        return permute_rows(x_own_global_order, self.own_order, self.own_inv)

    def to_global(self, x_own_local_order: torch.Tensor) -> torch.Tensor:
        from .processor import permute_rows
        if self.identity_order:
            return x_own_local_order
        return permute_rows(x_own_local_order, self.own_inv, self.own_order)

    def run(self, layers, x_own: torch.Tensor, e_csr: torch.Tensor):
        """x_own: own node rows in ascending global id ([hi - lo, D]); e_csr: local edges in the order of
        csr_edge_ids().  Returns (x', e') in the same orders.  The weight gradients the backward returns are already
        summed over the ranks (one all-reduce inside the stack's backward)."""
        layers = list(layers)
        cfg = layers[0].stack_config()
        flat = []
        for layer in layers:
            s = layer.step_weights(x_own.dtype)
            flat += [s.w_edge, s.w_node, s.w_proj, s.b_proj]
            self._stack_params.update(id(p) for p in layer.parameters())
        x, e = PartitionedStackFn.apply(cfg, self, self.to_local(x_own), e_csr, *flat)
        return self.to_global(x), e

    # ---- GMP / WeightedEdgeConv variant on the partition ------------------------------
    def extend(self, x_own_local: torch.Tensor) -> torch.Tensor:
        """[n_local, w]: own rows (LOCAL order) followed by the halo rows of their owners; differentiable."""
        return HaloExtendFn.apply(self, x_own_local.contiguous())

    def extend_static(self, t_own_global_order: torch.Tensor) -> torch.Tensor:
        """Per-node data that does not change during training (positions): exchanged ONCE per mesh, no gradient."""
        with torch.no_grad():
            loc = t_own_global_order[self.own_order.long()].contiguous()
            return HaloExtendFn.apply(self, loc)

    @property
    def local_edge_index(self) -> torch.Tensor:
        """[2, E_loc] int64 local (sender, receiver) ids in the order of halo.edge_ids (ascending global edge id)."""
        if not hasattr(self, "_local_ei"):
            self._local_ei = torch.from_numpy(self.halo.local_edge_index).to(self.plan.rowptr.device)
        return self._local_ei

    def weighted_edge_conv(self, conv, x_own: torch.Tensor, pos_ext: torch.Tensor, edge_weights=None,
                           compute_weights: bool = True):
        """WeightedEdgeConv.forward on this rank's block: x_own in ascending global id, pos_ext from extend_static().
        Returns (out for the own nodes in ascending global id, edge weights of the local edges in halo.edge_ids order).
        Every incoming edge of an own node is local, so the own rows are exact; gradients of halo rows travel back to
        their owners through HaloExtendFn; the module's weight gradients are per-rank partials (allreduce_grads)."""
        x_ext = self.extend(self.to_local(x_own))
        out, w = conv(x_ext, self.local_edge_index, pos_ext, edge_weights=edge_weights, compute_weights=compute_weights)
        return self.to_global(out[: self.n_own]), w

    def allreduce_grads(self, params) -> None:
        """Sum the gradients of parameters used OUTSIDE run() (encoders, decoder) over the ranks with one flat
        all-reduce; the gradients of the layers given to run() are summed inside the stack's backward already and
        are skipped here.  Every rank walks the SAME ordered parameter list and substitutes zeros for a missing
        gradient (a rank with an empty block has none), so the flat buffers always line up."""
        if self.world == 1:
            return
        ps = [p for p in params if id(p) not in self._stack_params]
        if not ps:
            return
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in ps])
        dist.all_reduce(flat, group=self.group)
        views, off = [], 0
        for p in ps:
            n = p.numel()
            if p.grad is None:
                p.grad = torch.empty_like(p)
            views.append(flat[off: off + n].view_as(p.grad))
            off += n
        torch._foreach_copy_([p.grad for p in ps], views)      # one multi-tensor kernel instead of one copy per parameter
