"""One large mesh across the GPUs of a box: contiguous receiver-node blocks + one halo exchange per step.

The reference has no distributed code; this is new.  Rank r owns nodes [lo, hi) and every edge
whose receiver it owns, so edge latents, the edge block and the receiver sums stay local and deterministic.  A
processor step reads only 1-hop sender latents (mgnLayer.py:40-41), so the only data crossing ranks is the latent
row of each remote sender ("halo"), once per step forward and the gradient of those rows once per step backward.
Weights are replicated; their gradients are summed with one all-reduce after the backward pass.

Local numbering: own nodes first (global id - lo), then halo nodes in ascending global id.  Because owners hold
contiguous id ranges, the halo rows coming from one peer are a contiguous slice of the halo block.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .processor import D, StackConfig


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
    send_idx: List[np.ndarray]      # own-local ids of the rows each peer needs, ascending global id

    @property
    def n_own(self) -> int:
        return self.hi - self.lo

    @property
    def n_halo(self) -> int:
        return int(self.halo_global.shape[0])

    @property
    def n_local(self) -> int:
        return self.n_own + self.n_halo


def build_halo_plan(edge_index: np.ndarray, n_nodes: int, rank: int, world: int) -> HaloPlan:
    """Every rank holds the full connectivity (16E bytes), so the plan needs no communication."""
    src, dst = edge_index[0], edge_index[1]
    lo, hi = block_bounds(n_nodes, world, rank)
    mine = np.flatnonzero((dst >= lo) & (dst < hi))
    s, d = src[mine], dst[mine] - lo
    remote = (s < lo) | (s >= hi)
    halo = np.unique(s[remote])
    loc = np.where(remote, (hi - lo) + np.searchsorted(halo, s), s - lo)
    recv_counts, send_idx = [], []
    for p in range(world):
        plo, phi = block_bounds(n_nodes, world, p)
        recv_counts.append(int(np.count_nonzero((halo >= plo) & (halo < phi))) if p != rank else 0)
        if p == rank:
            send_idx.append(np.empty(0, dtype=np.int64))
            continue
        theirs = (dst >= plo) & (dst < phi) & (src >= lo) & (src < hi)      # p's edges whose sender I own
        send_idx.append(np.unique(src[theirs]) - lo)
    return HaloPlan(rank, world, lo, hi, mine, np.stack([loc, d]).astype(np.int64), halo, recv_counts, send_idx)


class HaloExchanger:
    """Point-to-point halo exchange over torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, plan: HaloPlan, device, group=None):
        self.plan, self.group = plan, group
        self.send_idx = [torch.from_numpy(ix).to(device) for ix in plan.send_idx]
        self.recv_off = np.concatenate([[0], np.cumsum(plan.recv_counts)]).astype(int)

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

    # ---- split-phase API: post, do independent work on the compute stream, then finish ----------------------
    def forward_start(self, x_own: torch.Tensor, out: Optional[torch.Tensor] = None):
        """Post the exchange of the remote senders' rows into `out` ([n_halo, width], halo order; allocated when
        None).  Returns a token for forward_finish(); `out` is valid only after it."""
        pl = self.plan
        halo = out if out is not None else x_own.new_empty((pl.n_halo, x_own.size(1)))
        sends = [x_own[ix].contiguous() if ix.numel() else None for ix in self.send_idx]
        recvs = [halo[self.recv_off[p]: self.recv_off[p + 1]] if pl.recv_counts[p] else None for p in range(pl.world)]
        return halo, sends, self._post(sends, recvs)

    @staticmethod
    def forward_finish(token) -> torch.Tensor:
        halo, _sends, reqs = token
        for req in reqs:
            req.wait()
        return halo

    def backward_start(self, g_halo: torch.Tensor, like: torch.Tensor):
        """Post the return of the halo-row gradients to their owners (`like`: a tensor of the owners' dtype/device)."""
        pl = self.plan
        sends = [g_halo[self.recv_off[p]: self.recv_off[p + 1]].contiguous() if pl.recv_counts[p] else None
                 for p in range(pl.world)]
        recvs = [like.new_empty((ix.numel(), like.size(1))) if ix.numel() else None for ix in self.send_idx]
        return sends, recvs, self._post(sends, recvs)

    def backward_finish(self, token, g_own: torch.Tensor) -> None:
        """Owners add the returned rows in ascending peer order (deterministic: the rows one peer returns are
        distinct)."""
        _sends, recvs, reqs = token
        for req in reqs:
            req.wait()
        for p in range(self.plan.world):
            if recvs[p] is not None:
                g_own.index_add_(0, self.send_idx[p], recvs[p])

    # ---- blocking forms ------------------------------------------------------------------------------------
    def forward(self, x_own: torch.Tensor) -> torch.Tensor:
        """Rows of the remote senders, [n_halo, width], in halo order."""
        return self.forward_finish(self.forward_start(x_own))

    def backward(self, g_halo: torch.Tensor, g_own: torch.Tensor) -> None:
        """Send halo-row gradients back to their owners and accumulate them into g_own."""
        self.backward_finish(self.backward_start(g_halo, g_own), g_own)


class PartitionedStackFn(torch.autograd.Function):
    """MGN processor stack on one receiver block; apply(cfg, part, x_own, e_csr, *flat) like processor.MGNStackFn."""

    @staticmethod
    def forward(ctx, cfg: StackConfig, part: "PartitionedProcessor", x, e, *flat):
        plan, ex = part.plan, part.exchanger
        n_own = part.n_own
        K = len(flat) // 4
        x, e = x.contiguous(), e.contiguous()
        path_e = ops.choose_path(x.dtype, cfg.act_edge, cfg.L_edge)
        path_n = ops.choose_path(x.dtype, cfg.act_node, cfg.L_node)
        scale = plan.inv_deg[:n_own].contiguous() if cfg.mean else None
        paths_bwd = (ops.choose_path(x.dtype, cfg.act_edge, cfg.L_edge, backward=True),
                     ops.choose_path(x.dtype, cfg.act_node, cfg.L_node, backward=True))
        keep_h0 = ops.keeps_h0(path_e, path_n, *paths_bwd)
        saved, preps = [], []
        for k in range(K):
            w_edge, w_node, w_proj, b_proj = flat[4 * k: 4 * k + 4]
            pe = ops.PreparedBlock(w_edge.detach(), cfg.L_edge, path_e, cfg.act_edge, cfg.use_ln)
            pn = ops.PreparedBlock(w_node.detach(), cfg.L_node, path_n, cfg.act_node, cfg.use_ln)
            preps.append((pe, pn))
            # the halo rows travel while the own rows are copied and pre-projected (no dependence on the exchange)
            x_ext = x.new_empty((plan.N, D))
            tok = ex.forward_start(x, out=x_ext[n_own:])
            x_ext[:n_own].copy_(x)
            P = x.new_empty((plan.N, w_proj.size(0)))
            wt, bp = w_proj.detach().t(), b_proj.detach()
            torch.addmm(bp, x, wt, out=P[:n_own])
            ex.forward_finish(tok)
            if plan.N > n_own:
                torch.addmm(bp, x_ext[n_own:], wt, out=P[n_own:])
            h0e = torch.empty_like(e) if keep_h0 else None
            h0n = torch.empty_like(x) if keep_h0 else None
            e_new, agg = ops.block_fwd(pe, e, e, P, plan.src, plan.dst, 0, D, rowptr=plan.rowptr, want_agg=True,
                                       kind="edge_fwd", h0_out=h0e)
            agg = agg[:n_own]
            x_new, _ = ops.block_fwd(pn, agg, x, P, None, None, 2 * D, 0, main_scale=scale, kind="node_fwd",
                                     h0_out=h0n)
            # x_ext and (h_0 of both blocks | P) are kept so the backward needs no second halo exchange
            saved += [x_ext, e, agg, h0e, h0n] if keep_h0 else [x_ext, e, agg, P, P]
            x, e = x_new, e_new
        ctx.cfg, ctx.part, ctx.K = cfg, part, K
        ctx.set_materialize_grads(False)
        ctx.paths, ctx.keep_h0 = paths_bwd, keep_h0
        # the weight images of the forward serve the backward too when both run on the same kernel family
        ctx.preps = preps if (path_e, path_n) == paths_bwd else None
        ctx.save_for_backward(*saved, *flat)
        return x, e

    @staticmethod
    def backward(ctx, G_x, G_e):
        cfg, part, K = ctx.cfg, ctx.part, ctx.K
        plan, ex, n_own = part.plan, part.exchanger, part.n_own
        path_e, path_n = ctx.paths
        saved = ctx.saved_tensors
        acts, flat = saved[: 5 * K], saved[5 * K:]
        dt = acts[0].dtype
        G_x = torch.zeros_like(acts[0][:n_own]) if G_x is None else G_x.contiguous().to(dt)
        G_e = torch.zeros_like(acts[1]) if G_e is None else G_e.contiguous().to(dt).clone()
        scale = plan.inv_deg[:n_own].contiguous() if cfg.mean else None
        grads = [None] * (4 * K)
        for k in reversed(range(K)):
            x_ext, e, agg, a1, a2 = acts[5 * k: 5 * k + 5]
            P, h0e, h0n = (None, a1, a2) if ctx.keep_h0 else (a1, None, None)
            x = x_ext[:n_own]
            w_edge, w_node, w_proj, b_proj = flat[4 * k: 4 * k + 4]
            if ctx.preps is not None:
                pe, pn = ctx.preps[k]
            else:
                pe = ops.PreparedBlock(w_edge, cfg.L_edge, path_e, cfg.act_edge, cfg.use_ln)
                pn = ops.PreparedBlock(w_node, cfg.L_node, path_n, cfg.act_node, cfg.use_ln)
            g_agg, g_h0n, g_wn = ops.block_bwd(pn, agg, P, None, None, 2 * D, 0, G_x, main_scale=scale, kind="node_bwd",
                                               h0=h0n, n_nodes=plan.N)
            agg_eff = agg if scale is None else agg * scale[:, None]
            ops.wgrad_into(g_wn, g_h0n, agg_eff.to(dt))
            G_e, g_h0e, g_we = ops.block_bwd(pe, e, P, plan.src, plan.dst, 0, D, G_e, g_agg=g_agg, has_resid_grad=True,
                                             g_main_out=G_e, kind="edge_bwd", h0=h0e, n_nodes=plan.N,
                                             rowptr=plan.rowptr)
            ops.wgrad_into(g_we, g_h0e, e)
            g_psd = torch.empty((plan.N, 2 * D), dtype=dt, device=e.device)   # [g_P_s | g_P_d] over local rows
            ops.segment_reduce(g_h0e, plan.sptr, plan.sperm, plan.N, out=g_psd[:, :D])
            ops.segment_reduce(g_h0e, plan.rowptr, None, plan.N, out=g_psd[:, D:])
            g_ext = g_psd @ w_proj[:2 * D]                          # [n_local, D]
            tok = ex.backward_start(g_ext[n_own:], g_ext)           # halo-row gradients travel under the GEMMs below
            g_x = G_x + g_ext[:n_own]
            g_x.addmm_(g_h0n, w_proj[2 * D:])
            g_wproj = torch.empty_like(w_proj)
            torch.mm(g_psd.t(), x_ext, out=g_wproj[:2 * D])
            torch.mm(g_h0n.t(), x, out=g_wproj[2 * D:])
            g_bproj = torch.cat([g_we[-D:], g_we[-D:], g_wn[-D:]])
            ex.backward_finish(tok, g_x)
            grads[4 * k: 4 * k + 4] = [g_we, g_wn, g_wproj.to(w_proj.dtype), g_bproj.to(b_proj.dtype)]
            G_x = g_x
        return (None, None, G_x, G_e, *grads)


class PartitionedProcessor:
    """Receiver-block partition of one mesh for this rank: halo plan, local graph plan, exchange, grad all-reduce."""

    def __init__(self, edge_index: torch.Tensor, n_nodes: int, rank: int, world: int, device, group=None):
        ei = edge_index.cpu().numpy()
        self.halo = build_halo_plan(ei, n_nodes, rank, world)
        self.rank, self.world, self.group = rank, world, group
        self.lo, self.hi, self.n_own = self.halo.lo, self.halo.hi, self.halo.n_own
        self.edge_ids_cpu = torch.from_numpy(self.halo.edge_ids)
        local_ei = torch.from_numpy(self.halo.local_edge_index).to(device)
        self.plan = ops.build_graph_plan(local_ei, self.halo.n_local)
        self.E_loc = self.plan.E
        self.exchanger = HaloExchanger(self.halo, device, group)

    def csr_edge_ids(self) -> torch.Tensor:
        """Global caller edge id stored at each local CSR slot."""
        return self.edge_ids_cpu.to(self.plan.perm.device)[self.plan.perm.long()]

    def run(self, layers, x_own: torch.Tensor, e_csr: torch.Tensor):
        layers = list(layers)
        cfg = layers[0].stack_config()
        flat = []
        for layer in layers:
            s = layer.step_weights(x_own.dtype)
            flat += [s.w_edge, s.w_node, s.w_proj, s.b_proj]
        return PartitionedStackFn.apply(cfg, self, x_own, e_csr, *flat)

    def allreduce_grads(self, params) -> None:
        """Sum of the per-rank partial weight gradients (one flat all-reduce)."""
        ps = [p for p in params if p.grad is not None]
        if not ps or self.world == 1:
            return
        flat = torch.cat([p.grad.reshape(-1).float() for p in ps])
        dist.all_reduce(flat, group=self.group)
        views, off = [], 0
        for p in ps:
            n = p.numel()
            views.append(flat[off: off + n].view_as(p.grad))
            off += n
        torch._foreach_copy_([p.grad for p in ps], views)      # one multi-tensor kernel instead of one copy per parameter
