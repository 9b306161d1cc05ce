"""Bi-stride multi-scale MeshGraphNet (interface of reference models/bsms_mgn.py:9-306).

U-Net over mesh levels: processor layers per level, stride pooling on the way down (nodes of each
graph ordered by x, consecutive `stride` nodes merge; coarse edges are the distinct
(coarse sender, coarse receiver) pairs, self-loops kept), gather + skip on the way up with the
fine level's edge latents restored from the skip (bsms_mgn.py:199-206).

The level hierarchy (assignment, coarse connectivity, CSR plans) depends only on the mesh, so it is
built once by the integer kernels in aero_gnn_b200.pooling and cached; the reference rebuilds it on
every forward with host syncs per graph (bsms_mgn.py:231-262).
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from ..pooling import POOL_CACHE, PoolLevel, SegmentReduceFn, UnpoolAddFn
from ..processor import permute_rows
from ._common import encoder_kwargs, run_layers
from .mgnLayer import MeshGraphNetLayer
from .mlp import MLP


class BiStridedMeshGraphNet(nn.Module):
    def __init__(self, input_node_dim: int, input_edge_dim: int, output_node_dim: int, processor_size: int = 15,
                 activation_fn: str = "relu", num_hidden_layers_node_processor: int = 1,
                 num_hidden_layers_edge_processor: int = 1, hidden_dim_processor: int = 128,
                 num_hidden_layers_node_encoder: int = 1, hidden_dim_node_encoder: int = 128,
                 num_hidden_layers_edge_encoder: int = 1, hidden_dim_edge_encoder: int = 128,
                 aggregation: str = "add", hidden_dim_decoder: int = 128, num_hidden_layers_decoder: int = 1,
                 dropout: float = 0.0, do_concat_trick: bool = False, num_scales: int = 3, layers_per_scale=2,
                 stride: int = 2) -> None:
        super().__init__()
        if num_scales < 1:
            raise ValueError("num_scales must be >= 1")
        if stride < 1:
            raise ValueError("stride must be >= 1")
        self.num_scales, self.stride = num_scales, stride
        self.aggregation, self.do_concat_trick = aggregation, do_concat_trick
        H = hidden_dim_processor

        self.node_encoder = MLP(input_node_dim, hidden_dim_node_encoder, H, num_hidden_layers_node_encoder,
                                **encoder_kwargs(activation_fn, dropout))
        self.edge_encoder = MLP(input_edge_dim, hidden_dim_edge_encoder, H, num_hidden_layers_edge_encoder,
                                **encoder_kwargs(activation_fn, dropout))

        n_stage = max(num_scales - 1, 0)
        if isinstance(layers_per_scale, int):
            counts = [layers_per_scale] * n_stage
        else:
            if len(layers_per_scale) != n_stage:
                raise ValueError("layers_per_scale must be int or list with num_scales-1 elements")
            counts = list(layers_per_scale)
        n_bottleneck = max(1, processor_size - 2 * sum(counts))     # bsms_mgn.py:80-81

        def block(n: int) -> nn.ModuleList:
            return nn.ModuleList(
                MeshGraphNetLayer(H, H, H, num_hidden_layers_node_processor, num_hidden_layers_edge_processor,
                                  activation_fn, True, aggregation, do_concat_trick)
                for _ in range(n))

        self.down_layers = nn.ModuleList(block(c) for c in counts)
        self.bottleneck_layers = block(n_bottleneck)
        self.up_layers = nn.ModuleList(block(c) for c in reversed(counts))
        self.decoder = MLP(H, hidden_dim_decoder, output_node_dim, num_hidden_layers_decoder, activation_fn,
                           use_layer_norm=False)
        self.dropout = nn.Dropout(dropout) if dropout > 0 else None

    # ---- pooling -------------------------------------------------------------------------------
    def _downsample(self, node_attr, edge_attr, edge_index, batch, pos=None):
        """Reference-compatible signature and return tuple (bsms_mgn.py:217-301); edge_attr in caller order."""
        lvl = POOL_CACHE.get(edge_index, batch, pos, self.stride)
        coarse_x = SegmentReduceFn.apply(node_attr.contiguous(), lvl.node_gptr, lvl.node_glist, lvl.f2c32,
                                         lvl.n_coarse, True)
        coarse_pos = None
        if pos is not None:
            coarse_pos = ops.segment_reduce(pos.contiguous(), lvl.node_gptr, lvl.node_glist, lvl.n_coarse, mean=True)
        ec = lvl.coarse_edge_index.size(1)
        if ec > 0:
            coarse_e = SegmentReduceFn.apply(edge_attr.contiguous(), lvl.edge_gptr, lvl.edge_glist,
                                             lvl.inverse.to(torch.int32), ec, True)
        else:
            coarse_e = edge_attr.new_zeros((0, edge_attr.size(1) if edge_attr.dim() > 1 else 1))
        return coarse_x, coarse_e, lvl.coarse_edge_index, lvl.coarse_batch, coarse_pos, lvl.fine_to_coarse

    def _unpool_nodes(self, coarse_nodes: torch.Tensor, assignment: torch.Tensor) -> torch.Tensor:
        return ops.gather_rows(coarse_nodes, assignment.to(torch.int32))

    # ---- forward -------------------------------------------------------------------------------
    def forward(self, node_attr, edge_attr, edge_index, batch=None, pos=None):
        ops._require_cuda(node_attr, edge_attr, edge_index, batch, pos)
        if batch is None:
            batch = torch.zeros(node_attr.size(0), dtype=torch.long, device=node_attr.device)
        plan = ops.PLAN_CACHE.get(edge_index, node_attr.size(0))
        x = self.node_encoder(node_attr)
        e = self.edge_encoder(permute_rows(edge_attr, plan.perm, plan.inv_perm))   # CSR order from here on
        if self.dropout is not None:
            x, e = self.dropout(x), self.dropout(e)

        ei, b, p = edge_index, batch, pos
        skips = []
        for layers in self.down_layers:
            x, e = run_layers(layers, plan, x, e)
            lvl: PoolLevel = POOL_CACHE.get(ei, b, p, self.stride)
            skips.append((x, e, plan, lvl))
            cplan = ops.PLAN_CACHE.get(lvl.coarse_edge_index, lvl.n_coarse)
            # node latents / positions: mean over the fine nodes of each coarse node
            x = SegmentReduceFn.apply(x, lvl.node_gptr, lvl.node_glist, lvl.f2c32, lvl.n_coarse, True)
            if p is not None:
                # coarse positions are a function of the level's key (connectivity, batch, positions): computed once
                # per level, so the next level's lookup sees the same tensor object (identity fast path, no hashing)
                cp = lvl.__dict__.get("_coarse_pos")
                if cp is None:
                    cp = ops.segment_reduce(p.contiguous(), lvl.node_gptr, lvl.node_glist, lvl.n_coarse, mean=True)
                    lvl.__dict__["_coarse_pos"] = cp
                p = cp
            # edge latents: mean over the fine edges of each coarse edge, straight into coarse CSR order
            e = _pool_edges(e, plan, lvl, cplan)
            ei, b, plan = lvl.coarse_edge_index, lvl.coarse_batch, cplan

        x, e = run_layers(self.bottleneck_layers, plan, x, e)

        for i, layers in enumerate(self.up_layers):
            sx, se, splan, lvl = skips[-(i + 1)]
            x = UnpoolAddFn.apply(x, sx, lvl.f2c32, lvl.node_gptr, lvl.node_glist)
            e, plan = se, splan
            x, e = run_layers(layers, plan, x, e)
        return self.decoder(x)


def _pool_edges(e_csr: torch.Tensor, plan: ops.GraphPlan, lvl: PoolLevel, cplan: ops.GraphPlan) -> torch.Tensor:
    """Mean-pool fine edge latents (fine CSR order) into coarse edge latents (coarse CSR order).

    Members of a coarse edge are visited in ascending fine caller edge id -- the accumulation order of the
    reference's CPU scatter (bsms_mgn.py:283)."""
    ec = lvl.coarse_edge_index.size(1)
    if ec == 0:
        return e_csr.new_zeros((0, e_csr.size(1)))
    cache = lvl.__dict__.setdefault("_edge_pool_cache", {})
    key = id(plan)
    if key not in cache or cache[key][0] is not plan:
        slots = plan.inv_perm[lvl.edge_glist.long()].contiguous()                  # fine CSR slot of each member
        group_of_slot = lvl.inverse[plan.perm.long()].to(torch.int32).contiguous()  # coarse edge of each fine slot
        cache[key] = (plan, slots, group_of_slot)
    _, slots, group_of_slot = cache[key]
    e_u = SegmentReduceFn.apply(e_csr, lvl.edge_gptr, slots, group_of_slot, ec, True)  # unique-key order
    return permute_rows(e_u, cplan.perm, cplan.inv_perm)
