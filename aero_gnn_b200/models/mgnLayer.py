"""One MeshGraphNets processor step (interface of reference models/mgnLayer.py:10-213).

Same classes, constructor arguments, forward signatures and state_dict keys as the reference;
the arithmetic runs in the fused sm_100a block kernels (aero_gnn_b200.processor).  The reference's
per-step cuda.synchronize()/allocator prints (mgnLayer.py:186-203) are instrumentation, not
behaviour, and are not reproduced.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from ..processor import (D, StackConfig, StepWeights, cached_pack_step, run_stack, permute_rows, edge_block_apply,
                         node_block_apply)
from .mlp import MLP


def _check_width(node_dim: int, edge_dim: int, hidden_dim: int) -> None:
    if not (node_dim == edge_dim == hidden_dim == D):
        raise RuntimeError(
            f"the fused sm_100a path supports node_dim == edge_dim == hidden_dim == {D} only "
            f"(got {node_dim}, {edge_dim}, {hidden_dim})"
        )


class EdgeBlock(nn.Module):
    """Edge update MLP over cat[e, x[sender], x[receiver]] (reference mgnLayer.py:10-49)."""

    def __init__(self, node_dim: int, edge_dim: int, hidden_dim: int = 128, num_hidden_layers: int = 1,
                 activation_fn: str = "relu", use_layer_norm: bool = True):
        super().__init__()
        self.dims = (node_dim, edge_dim, hidden_dim)
        self.mlp = MLP(edge_dim + 2 * node_dim, hidden_dim, edge_dim, num_hidden_layers, activation_fn,
                       use_layer_norm=use_layer_norm)

    # first Linear split by input block: [W_e | W_s | W_d]
    def fused_parts(self):
        _check_width(*self.dims)
        if self.mlp.num_hidden < 0:
            raise RuntimeError("EdgeBlock with num_hidden_layers=0 (a single Linear) has no fused sm_100a form")
        nd, ed, _ = self.dims
        w0, b0 = self.mlp.layers[0].weight, self.mlp.layers[0].bias
        hidden, w_out, b_out, gamma, beta = self.mlp.tail()
        return dict(w_e=w0[:, :ed], w_s=w0[:, ed:ed + nd], w_d=w0[:, ed + nd:], b0=b0, hidden=hidden,
                    w_out=w_out, b_out=b_out, gamma=gamma, beta=beta, act=self.mlp.activation_name,
                    use_ln=self.mlp.use_layer_norm)

    def forward(self, edge_attr, node_attr, edge_index):
        return edge_block_apply(self.fused_parts(), edge_attr, node_attr, edge_index)


class EdgeBlockSum(nn.Module):
    """Sum-trick edge block (reference mgnLayer.py:51-105): first Linear held as three [H, D] parameters
    `edge_lin`, `src_lin`, `dst_lin` + `bias`; ReLU regardless of `activation_fn` (mgnLayer.py:81)."""

    def __init__(self, node_dim: int, edge_dim: int, hidden_dim: int = 128, num_hidden_layers: int = 1,
                 activation_fn: str = "relu", use_layer_norm: bool = True):
        super().__init__()
        self.dims = (node_dim, edge_dim, hidden_dim)
        self.edge_dim, self.src_dim, self.dst_dim = edge_dim, node_dim, node_dim
        self.num_hidden_layers = num_hidden_layers
        self.use_layer_norm = use_layer_norm
        first = nn.Linear(edge_dim + 2 * node_dim, hidden_dim)  # same init stream as the reference
        w_e, w_s, w_d = first.weight.detach().split([edge_dim, node_dim, node_dim], dim=1)
        self.edge_lin = nn.Parameter(w_e.clone())
        self.src_lin = nn.Parameter(w_s.clone())
        self.dst_lin = nn.Parameter(w_d.clone())
        self.bias = nn.Parameter(first.bias.detach().clone())
        relu = nn.ReLU()
        mods = [relu]
        for _ in range(num_hidden_layers):
            mods += [nn.Linear(hidden_dim, hidden_dim), relu]
        mods.append(nn.Linear(hidden_dim, edge_dim))
        if use_layer_norm:
            mods.append(nn.LayerNorm(edge_dim))
        self.mlp = nn.Sequential(*mods)   # keys mlp.{1,3,..}.weight|bias, mlp.{2L+2} = LayerNorm

    def fused_parts(self):
        _check_width(*self.dims)
        L = self.num_hidden_layers
        hidden = [(self.mlp[1 + 2 * l].weight, self.mlp[1 + 2 * l].bias) for l in range(L)]
        out = self.mlp[1 + 2 * L]
        if self.use_layer_norm:
            gamma, beta = self.mlp[2 + 2 * L].weight, self.mlp[2 + 2 * L].bias
        else:
            gamma = torch.ones_like(out.bias)
            beta = torch.zeros_like(out.bias)
        return dict(w_e=self.edge_lin, w_s=self.src_lin, w_d=self.dst_lin, b0=self.bias, hidden=hidden,
                    w_out=out.weight, b_out=out.bias, gamma=gamma, beta=beta, act="relu",
                    use_ln=self.use_layer_norm)

    def forward(self, edge_attr, node_attr, edge_index):
        return edge_block_apply(self.fused_parts(), edge_attr, node_attr, edge_index)


class NodeBlock(nn.Module):
    """Node update MLP over cat[x, aggregate of incoming e] (reference mgnLayer.py:111-153)."""

    def __init__(self, node_dim: int, edge_dim: int, hidden_dim: int = 128, num_hidden_layers: int = 1,
                 activation_fn: str = "relu", use_layer_norm: bool = True, aggregation: str = "add"):
        super().__init__()
        self.dims = (node_dim, edge_dim, hidden_dim)
        self.aggregation = aggregation
        self.mlp = MLP(node_dim + edge_dim, hidden_dim, node_dim, num_hidden_layers, activation_fn,
                       use_layer_norm=use_layer_norm)

    def check_aggregation(self) -> bool:
        """True for 'mean', False for 'add'; anything else raises like mgnLayer.py:147-148."""
        if self.aggregation == "mean":
            return True
        if self.aggregation == "add":
            return False
        raise ValueError(f"Unsupported aggregation method: {self.aggregation}")

    def fused_parts(self):
        _check_width(*self.dims)
        if self.mlp.num_hidden < 0:
            raise RuntimeError("NodeBlock with num_hidden_layers=0 (a single Linear) has no fused sm_100a form")
        nd = self.dims[0]
        w0, b0 = self.mlp.layers[0].weight, self.mlp.layers[0].bias
        hidden, w_out, b_out, gamma, beta = self.mlp.tail()
        return dict(w_x=w0[:, :nd], w_a=w0[:, nd:], b0=b0, hidden=hidden, w_out=w_out, b_out=b_out, gamma=gamma,
                    beta=beta, act=self.mlp.activation_name, use_ln=self.mlp.use_layer_norm)

    def forward(self, node_attr, edge_attr, edge_index):
        mean = self.check_aggregation()
        return node_block_apply(self.fused_parts(), node_attr, edge_attr, edge_index, mean)


class MeshGraphNetLayer(nn.Module):
    """Edge block + node block with both residuals (reference mgnLayer.py:156-213)."""

    def __init__(self, node_dim: int, edge_dim: int, hidden_dim: int = 128,
                 num_hidden_layers_node_processor: int = 1, num_hidden_layers_edge_processor: int = 1,
                 activation_fn: str = "relu", use_layer_norm: bool = True, aggregation: str = "add",
                 do_concat_trick: bool = False):
        super().__init__()
        edge_cls = EdgeBlockSum if do_concat_trick else EdgeBlock
        self.edge_block = edge_cls(node_dim, edge_dim, hidden_dim, num_hidden_layers_edge_processor, activation_fn,
                                   use_layer_norm)
        self.node_block = NodeBlock(node_dim, edge_dim, hidden_dim, num_hidden_layers_node_processor, activation_fn,
                                    use_layer_norm, aggregation)

    # ---- fused-path views -------------------------------------------------------------------------
    def stack_config(self) -> StackConfig:
        ep, np_ = self.edge_block.fused_parts(), self.node_block.fused_parts()
        if ep["use_ln"] != np_["use_ln"]:
            raise RuntimeError("edge and node blocks must agree on use_layer_norm")
        return StackConfig(L_edge=len(ep["hidden"]), L_node=len(np_["hidden"]), act_edge=ep["act"],
                           act_node=np_["act"], use_ln=ep["use_ln"], mean=self.node_block.check_aggregation())

    def step_weights(self, dtype: torch.dtype) -> StepWeights:
        def build():
            ep, np_ = self.edge_block.fused_parts(), self.node_block.fused_parts()
            return ((ep["w_e"], ep["hidden"], ep["w_out"], ep["b_out"], ep["gamma"], ep["beta"]),
                    (np_["w_a"], np_["hidden"], np_["w_out"], np_["b_out"], np_["gamma"], np_["beta"]),
                    [ep["w_s"], ep["w_d"], np_["w_x"]], [None, ep["b0"], np_["b0"]])
        return cached_pack_step(self, dtype, build)

    def forward(self, node_attr, edge_attr, edge_index):
        ops._require_cuda(node_attr, edge_attr, edge_index)
        cfg = self.stack_config()
        plan = ops.PLAN_CACHE.get(edge_index, node_attr.size(0))
        e_csr = permute_rows(edge_attr, plan.perm, plan.inv_perm)
        x, e_csr = run_stack(cfg, plan, node_attr, e_csr, [self.step_weights(node_attr.dtype)])
        return x, permute_rows(e_csr, plan.inv_perm, plan.perm)
