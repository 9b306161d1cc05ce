"""BSMS-GNN variant with BFS-bistride pooling (interface of the reference's older `models.bsms_mgn`, which survives
only as bytecode: models/__pycache__/bsms_mgn.cpython-311.pyc; "orig :NN" = first line of a code object in it;
behavioural spec in SURVEY.md section 2.3).  Classes: MultiScaleGraphPreprocessor (orig :18), BSMSGMP (orig :96),
BSMS_MeshGraphNet (orig :203), create_bsms_model_from_config (orig :340).

The hierarchy (BFS levels, selected nodes, coarse edge lists) is integer work done by the kernels of
csrc/bistride.cu once per mesh and cached by content hash; message passing runs on the fused block kernels (GMP) and
the WeightedEdgeConv gather kernels.  CUDA tensors only.
"""
from __future__ import annotations

from typing import Dict, List

import torch
from torch import nn

from .. import bistride as _b
from .. import ops
from ..processor import permute_rows
from .bistride_ops import GMP, Unpool, WeightedEdgeConv, _index_maps
from .mlp import MLP


class MultiScaleGraphPreprocessor:
    """Builds the level hierarchy of one mesh (orig :18-94)."""

    def __init__(self, num_levels: int = 3):
        self.num_levels = num_levels

    def create_multiscale_graph(self, data) -> Dict[str, List]:
        """data.edge_index [2,E], data.pos [N,p] -> dict of lists: edge_indices (levels 0..L), node_indices (L),
        num_nodes (L+1), positions (L+1).  Per level: even-BFS-distance nodes are kept (orig :56), edges with both
        endpoints kept are renumbered, self-loops dropped (orig :32)."""
        edge_index, pos = data.edge_index, data.pos
        ops._require_cuda(edge_index, pos)
        n = int(pos.shape[0])
        multi = {"edge_indices": [edge_index], "node_indices": [], "num_nodes": [n], "positions": [pos]}
        for _ in range(self.num_levels):
            lvl = _b.BISTRIDE_CACHE.get(edge_index, n, pos)
            multi["node_indices"].append(lvl.selected)
            pos = pos[lvl.selected]
            edge_index = lvl.coarse_edge_index
            n = int(lvl.selected.numel())
            multi["positions"].append(pos)
            multi["edge_indices"].append(edge_index)
            multi["num_nodes"].append(n)
        return multi


class BSMSGMP(nn.Module):
    """Bistride multi-scale message passing (orig :96-201): GMP + WeightedEdgeConv on the way down, row-subset pooling,
    a bottom GMP, then unpool + WeightedEdgeConv with the down pass's edge weights + skip on the way up."""

    def __init__(self, num_levels: int, latent_dim: int, hidden_dim: int, pos_dim: int = 2):
        super().__init__()
        self.num_levels, self.latent_dim, self.pos_dim = num_levels, latent_dim, pos_dim
        self.down_gmps = nn.ModuleList(GMP(latent_dim, latent_dim, hidden_dim) for _ in range(num_levels + 1))
        self.down_edge_convs = nn.ModuleList(WeightedEdgeConv(latent_dim, latent_dim, aggr="add")
                                             for _ in range(num_levels))
        self.bottom_gmp = GMP(latent_dim, latent_dim, hidden_dim)
        self.up_edge_convs = nn.ModuleList(WeightedEdgeConv(latent_dim, latent_dim, aggr="add")
                                           for _ in range(num_levels))
        self.unpools = nn.ModuleList(Unpool() for _ in range(num_levels))

    def _gmp(self, gmp: GMP, x, edge_attr, edge_index, in_csr_order: bool):
        """One GMP step; the updated edge latents are dropped -- nothing downstream reads them (orig :145 returns x
        only), so they are neither permuted back to the caller's edge order nor kept."""
        ops._require_cuda(x, edge_attr, edge_index)
        plan = ops.PLAN_CACHE.get(edge_index, x.size(0))
        e_csr = edge_attr if in_csr_order else permute_rows(edge_attr, plan.perm, plan.inv_perm)
        return gmp.forward_csr(x, e_csr, plan)[0]

    def forward(self, x, edge_attrs, edge_indices, node_indices, num_nodes_list, positions,
                edges_in_csr_order: bool = False):
        """`edges_in_csr_order` (extension, used by BSMS_MeshGraphNet): edge_attrs[i] rows are already in the
        receiver-CSR order of ops.PLAN_CACHE.get(edge_indices[i], n_i) instead of the caller's edge order."""
        skips, weights_down = [], []
        for i in range(self.num_levels):
            x = self._gmp(self.down_gmps[i], x, edge_attrs[i], edge_indices[i], edges_in_csr_order)
            skips.append(x)                                    # the reference clones; nothing here writes in place
            x_conv, ew = self.down_edge_convs[i](x, edge_indices[i], positions[i], compute_weights=True)
            weights_down.append(ew)
            x = x + x_conv
            sel32, imap = _index_maps(node_indices[i], int(num_nodes_list[i]))
            x = _b.SelectRowsFn.apply(x, sel32, imap)          # x[node_indices[i]]
        x = self._gmp(self.bottom_gmp, x, edge_attrs[-1], edge_indices[-1], edges_in_csr_order)
        for i in range(self.num_levels - 1, -1, -1):
            x = self.unpools[i](x, node_indices[i], num_nodes_list[i])
            x_conv, _ = self.up_edge_convs[i](x, edge_indices[i], positions[i], edge_weights=weights_down[i],
                                              compute_weights=False)
            x = x + x_conv + skips[i]
        return x


class BSMS_MeshGraphNet(nn.Module):
    """Encoder -> BSMSGMP -> decoder (orig :203-337).  forward needs the hierarchy from MultiScaleGraphPreprocessor."""

    def __init__(self, input_node_dim: int, input_edge_dim: int, output_node_dim: int, num_levels: int = 3,
                 latent_dim: int = 128, hidden_dim: int = 128, pos_dim: int = 2, num_hidden_layers_encoder: int = 2,
                 num_hidden_layers_decoder: int = 2, activation_fn: str = "relu", dropout: float = 0.0):
        super().__init__()
        self.num_levels, self.latent_dim, self.pos_dim = num_levels, latent_dim, pos_dim
        self.node_encoder = MLP(input_dim=input_node_dim, hidden_dim=hidden_dim, output_dim=latent_dim,
                                num_hidden_layers=num_hidden_layers_encoder, activation_fn=activation_fn,
                                dropout=dropout, use_layer_norm=True)
        self.edge_encoder = MLP(input_dim=input_edge_dim, hidden_dim=hidden_dim, output_dim=latent_dim,
                                num_hidden_layers=num_hidden_layers_encoder, activation_fn=activation_fn,
                                dropout=dropout, use_layer_norm=True)
        self.bsgmp = BSMSGMP(num_levels=num_levels, latent_dim=latent_dim, hidden_dim=hidden_dim, pos_dim=pos_dim)
        self.decoder = MLP(input_dim=latent_dim, hidden_dim=hidden_dim, output_dim=output_node_dim,
                           num_hidden_layers=num_hidden_layers_decoder, activation_fn=activation_fn, dropout=dropout,
                           use_layer_norm=False)

    def forward(self, node_attr, edge_attr, edge_index, multi_data=None):
        ops._require_cuda(node_attr, edge_attr, edge_index)
        if multi_data is None:
            raise ValueError("multi_data must be provided. Use MultiScaleGraphPreprocessor to preprocess graphs before training.")
        node_hidden = self.node_encoder(node_attr)
        # raw edge features go into receiver-CSR order once, before the encoder (a few columns per edge instead of a
        # 128-wide latent row per GMP call); coarse-level edge latents start as zeros, which any order leaves unchanged
        plan0 = ops.PLAN_CACHE.get(multi_data["edge_indices"][0], node_attr.size(0))
        edge_hidden = self.edge_encoder(permute_rows(edge_attr, plan0.perm, plan0.inv_perm))
        edge_attrs = [edge_hidden]
        for i in range(1, len(multi_data["edge_indices"])):
            num_edges = multi_data["edge_indices"][i].shape[1]
            edge_attrs.append(torch.zeros(num_edges, self.latent_dim, device=edge_hidden.device, dtype=edge_hidden.dtype))
        x = self.bsgmp(node_hidden, edge_attrs, multi_data["edge_indices"], multi_data["node_indices"],
                       multi_data["num_nodes"], multi_data["positions"], edges_in_csr_order=True)
        return self.decoder(x)


def create_bsms_model_from_config(config: dict) -> BSMS_MeshGraphNet:
    """BSMS_MeshGraphNet from config['model'] with the reference's defaults (orig :340)."""
    m = config["model"]
    return BSMS_MeshGraphNet(input_node_dim=m["input_node_dim"], input_edge_dim=m["input_edge_dim"],
                             output_node_dim=m["output_node_dim"], num_levels=m.get("num_levels", 3),
                             latent_dim=m.get("hidden_dim", 128), hidden_dim=m.get("hidden_dim", 128),
                             pos_dim=m.get("pos_dim", 2),
                             num_hidden_layers_encoder=m.get("num_hidden_layers_encoder", 2),
                             num_hidden_layers_decoder=m.get("num_hidden_layers_decoder", 2),
                             activation_fn=m.get("activation_fn", "relu"), dropout=m.get("dropout", 0.0))
