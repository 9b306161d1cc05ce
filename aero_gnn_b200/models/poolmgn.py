"""MeshGraphNet with a pooled global feature broadcast to every node before the node encoder
(interface of reference models/poolmgn.py:11-157)."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from ..pooling import SegmentReduceFn, group_lists
from ..processor import permute_rows
from ._common import encoder_kwargs, run_layers
from .mgnLayer import MeshGraphNetLayer
from .mlp import MLP


class poolMGN(nn.Module):
    def __init__(self, input_node_dim: int, input_edge_dim: int, output_node_dim: int, processor_size: int = 15,
                 activation_fn: str = "relu", num_hidden_layers_node_processor: int = 1,
                 num_hidden_layers_edge_processor: int = 1, hidden_dim_processor: int = 128,
                 num_hidden_layers_node_encoder: int = 1, hidden_dim_node_encoder: int = 128,
                 num_hidden_layers_edge_encoder: int = 1, hidden_dim_edge_encoder: int = 128,
                 aggregation: str = "sum", hidden_dim_decoder: int = 128, num_hidden_layers_decoder: int = 1,
                 global_pool_method: str = "mean", num_hidden_layers_global_encoder: int = 1, global_dim: int = 128,
                 dropout: float = 0.0):
        super().__init__()
        if global_pool_method not in ("mean", "max", "add"):
            raise ValueError(f"Unsupported global pooling method: {global_pool_method}")
        self.global_pool_method = global_pool_method
        H = hidden_dim_processor
        self.node_encoder = MLP(input_node_dim + global_dim, hidden_dim_node_encoder, H, num_hidden_layers_node_encoder,
                                **encoder_kwargs(activation_fn, dropout))
        self.edge_encoder = MLP(input_edge_dim, hidden_dim_edge_encoder, H, num_hidden_layers_edge_encoder,
                                **encoder_kwargs(activation_fn, dropout))
        self.global_encoder = MLP(input_node_dim, global_dim, global_dim, num_hidden_layers_global_encoder,
                                  activation_fn, dropout=dropout, use_layer_norm=False)
        self.layers = nn.ModuleList(
            MeshGraphNetLayer(H, H, H, num_hidden_layers_node_processor, num_hidden_layers_edge_processor,
                              activation_fn, True, aggregation)
            for _ in range(processor_size))
        self.decoder = MLP(H, hidden_dim_decoder, output_node_dim, num_hidden_layers_decoder, activation_fn,
                           use_layer_norm=False)

    def global_pool(self, feats: torch.Tensor, batch: torch.Tensor) -> torch.Tensor:
        """Per-graph reduce of node rows (torch_geometric global_{mean,max,add}_pool, poolmgn.py:38-42)."""
        n_graphs = int(batch.max().item()) + 1 if batch.numel() else 0
        if self.global_pool_method == "max":
            out = feats.new_full((n_graphs, feats.size(1)), float("-inf"))
            idx = batch.view(-1, 1).expand_as(feats)
            return out.scatter_reduce(0, idx, feats, reduce="amax", include_self=True)
        gptr, glist, g32 = group_lists(batch, n_graphs)
        return SegmentReduceFn.apply(feats.contiguous(), gptr, glist, g32, n_graphs, self.global_pool_method == "mean")

    def forward(self, node_attr, edge_attr, edge_index, batch=None):
        ops._require_cuda(node_attr, edge_attr, edge_index, batch)
        g = self.global_encoder(node_attr)
        if batch is None:
            batch = torch.zeros(node_attr.size(0), dtype=torch.long, device=node_attr.device)
        pooled = self.global_pool(g, batch)
        g = pooled[batch]          # == repeat_interleave(bincount(batch)) for sorted batch (poolmgn.py:135)
        plan = ops.PLAN_CACHE.get(edge_index, node_attr.size(0))
        x = self.node_encoder(torch.cat((node_attr, g), dim=-1))
        e = self.edge_encoder(permute_rows(edge_attr, plan.perm, plan.inv_perm))
        x, _ = run_layers(self.layers, plan, x, e)
        return self.decoder(x)
