"""MeshGraphNet with Fourier positional features on the leading node inputs
(interface of reference models/fouriermgn.py:10-183)."""
from __future__ import annotations

import math

import torch
from torch import nn

from .. import ops
from ..processor import permute_rows
from ._common import encoder_kwargs, run_layers
from .mgnLayer import MeshGraphNetLayer
from .mlp import MLP


class FourierMeshGraphNet(nn.Module):
    def __init__(self, input_node_dim: int, input_edge_dim: int, output_node_dim: int, processor_size: int = 15,
                 activation_fn: str = "relu", num_hidden_layers_node_processor: int = 1,
                 num_hidden_layers_edge_processor: int = 1, hidden_dim_processor: int = 128,
                 num_hidden_layers_node_encoder: int = 1, hidden_dim_node_encoder: int = 128,
                 num_hidden_layers_edge_encoder: int = 1, hidden_dim_edge_encoder: int = 128,
                 aggregation: str = "sum", hidden_dim_decoder: int = 128, num_hidden_layers_decoder: int = 1,
                 dropout: float = 0.0, fourier_features_dim: int = 2, fourier_freq_start: int = -3,
                 fourier_freq_length: int = 7):
        super().__init__()
        self.fourier_features_dim = fourier_features_dim
        self.fourier_freq_start = fourier_freq_start
        self.fourier_freq_length = fourier_freq_length
        H = hidden_dim_processor
        expanded = input_node_dim + 2 * fourier_freq_length * fourier_features_dim
        self.node_encoder = MLP(expanded, hidden_dim_node_encoder, H, num_hidden_layers_node_encoder,
                                **encoder_kwargs(activation_fn, dropout))
        self.edge_encoder = MLP(input_edge_dim, hidden_dim_edge_encoder, H, num_hidden_layers_edge_encoder,
                                **encoder_kwargs(activation_fn, dropout))
        self.layers = nn.ModuleList(
            MeshGraphNetLayer(H, H, H, num_hidden_layers_node_processor, num_hidden_layers_edge_processor,
                              activation_fn, True, aggregation)
            for _ in range(processor_size))
        self.decoder = MLP(H, hidden_dim_decoder, output_node_dim, num_hidden_layers_decoder, activation_fn,
                           use_layer_norm=False)

    def fourier_embedding(self, pos: torch.Tensor) -> torch.Tensor:
        """[cos(2^i pi x) for all i | sin(2^i pi x) for all i] per spatial dim, flattened per node
        (layout of fouriermgn.py:143-149: [N, dim, 2*F] -> [N, dim*2*F])."""
        # phases and sin/cos are evaluated in (at least) fp32 and rounded once: with bf16 latents the reference's
        # bf16-mode phase 2^i*pi*x (up to ~25 rad) would carry ~0.05 rad of rounding error
        wide = torch.float32 if pos.dtype in (torch.bfloat16, torch.float16) else pos.dtype
        xs = pos[:, : self.fourier_features_dim].to(wide)
        k = torch.arange(self.fourier_freq_start, self.fourier_freq_start + self.fourier_freq_length,
                         device=pos.device, dtype=wide)
        phase = ((2.0 ** k) * math.pi).view(1, 1, -1) * xs.unsqueeze(-1)
        emb = torch.cat([torch.cos(phase), torch.sin(phase)], dim=-1).reshape(pos.shape[0], -1)
        return emb.to(pos.dtype)

    def forward(self, node_attr, edge_attr, edge_index):
        ops._require_cuda(node_attr, edge_attr, edge_index)
        plan = ops.PLAN_CACHE.get(edge_index, node_attr.size(0))
        x = self.node_encoder(torch.cat([node_attr, self.fourier_embedding(node_attr)], dim=-1))
        e = self.edge_encoder(permute_rows(edge_attr, plan.perm, plan.inv_perm))
        x, _ = run_layers(self.layers, plan, x, e)
        return self.decoder(x)
