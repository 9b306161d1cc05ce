"""Encode - process - decode MeshGraphNet (interface of reference models/mgn.py:9-139)."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from ..processor import permute_rows
from ._common import encoder_kwargs, run_layers
from .mgnLayer import MeshGraphNetLayer
from .mlp import MLP


class MeshGraphNet(nn.Module):
    """Same constructor keywords and state_dict keys (`node_encoder.*`, `edge_encoder.*`, `layers.{i}.*`,
    `decoder.*`) as the reference.  Note the reference default aggregation='sum' is rejected by NodeBlock
    at the first forward (mgnLayer.py:147-148); config.yaml:44 passes 'add'."""

    def __init__(self, input_node_dim: int, input_edge_dim: int, output_node_dim: int, processor_size: int = 15,
                 activation_fn: str = "relu", num_hidden_layers_node_processor: int = 1,
                 num_hidden_layers_edge_processor: int = 1, hidden_dim_processor: int = 128,
                 num_hidden_layers_node_encoder: int = 1, hidden_dim_node_encoder: int = 128,
                 num_hidden_layers_edge_encoder: int = 1, hidden_dim_edge_encoder: int = 128,
                 aggregation: str = "sum", hidden_dim_decoder: int = 128, num_hidden_layers_decoder: int = 1,
                 dropout: float = 0.0, do_concat_trick: bool = False):
        super().__init__()
        H = hidden_dim_processor
        self.node_encoder = MLP(input_node_dim, hidden_dim_node_encoder, H, num_hidden_layers_node_encoder,
                                **encoder_kwargs(activation_fn, dropout))
        self.edge_encoder = MLP(input_edge_dim, hidden_dim_edge_encoder, H, num_hidden_layers_edge_encoder,
                                **encoder_kwargs(activation_fn, dropout))
        self.layers = nn.ModuleList(
            MeshGraphNetLayer(H, H, H, num_hidden_layers_node_processor, num_hidden_layers_edge_processor,
                              activation_fn, True, aggregation, do_concat_trick)
            for _ in range(processor_size))
        self.decoder = MLP(H, hidden_dim_decoder, output_node_dim, num_hidden_layers_decoder, activation_fn,
                           use_layer_norm=False)

    def forward(self, node_attr: torch.Tensor, edge_attr: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        ops._require_cuda(node_attr, edge_attr, edge_index)
        plan = ops.PLAN_CACHE.get(edge_index, node_attr.size(0))
        # raw edge features go to receiver-CSR order once; no model returns edge latents (mgn.py:130)
        edge_csr = permute_rows(edge_attr, plan.perm, plan.inv_perm)
        x = self.node_encoder(node_attr)
        e = self.edge_encoder(edge_csr)
        x, _ = run_layers(self.layers, plan, x, e)
        return self.decoder(x)
