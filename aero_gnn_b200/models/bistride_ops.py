"""BFS-bistride operators (interface of the reference's `models.bistride_ops`, which ships only as bytecode:
models/__pycache__/bistride_ops.cpython-311.pyc; "orig :NN" = first line of a code object in it; behavioural spec
in SURVEY.md section 2.3).

Same class names, constructor arguments, forward signatures and state_dict keys; the arithmetic runs in the sm_100a
kernels behind aero_gnn_b200.bistride (BFS levels, selection, WeightedEdgeConv) and aero_gnn_b200.processor (GMP =
one fused message-passing step with two-Linear MLPs).  CUDA tensors only.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .. import bistride as _b
from .. import ops
from ..processor import D, StackConfig, StepWeights, cached_pack_step, permute_rows, run_stack


class BistridePooling:
    """Node selection on every other BFS frontier (orig :13-94)."""

    @staticmethod
    def bfs_distance(edge_index: torch.Tensor, num_nodes: int, start_node: int) -> torch.Tensor:
        """int64 [num_nodes] hop counts from start_node along edge_index[0] -> edge_index[1]; -1 = unreachable
        (orig :21: a Python deque loop with one .item() per edge; here one kernel launch per BFS level)."""
        ops._require_cuda(edge_index)
        plan = ops.PLAN_CACHE.get(edge_index, num_nodes)
        return _b.bfs_levels(plan, int(start_node))

    @staticmethod
    def select_bistride_nodes(edge_index: torch.Tensor, num_nodes: int, pos: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Ascending int64 ids of the nodes at even BFS distance from the seed (orig :56); every reached node when
        that is fewer than 30 % of the nodes."""
        return _b.select_bistride_nodes(edge_index, num_nodes, pos)


_IMAP_CACHE: dict = {}


def _index_maps(indices: torch.Tensor, n_fine: int):
    """(indices as int32, inverse map int32 [n_fine] with -1 for rows that were not kept); remembered per live
    `indices` tensor, since a hierarchy's node_indices are reused by every forward."""
    key = (indices.data_ptr(), int(indices.numel()), int(n_fine), indices._version, str(indices.device))
    hit = _IMAP_CACHE.get(key)
    if hit is not None and hit[0]() is indices:
        return hit[1], hit[2]
    sel32 = indices.to(torch.int32)
    imap = torch.full((n_fine,), -1, dtype=torch.int32, device=indices.device)
    imap[indices.long()] = torch.arange(indices.numel(), dtype=torch.int32, device=indices.device)
    if len(_IMAP_CACHE) > 64:
        _IMAP_CACHE.clear()
    import weakref
    _IMAP_CACHE[key] = (weakref.ref(indices), sel32, imap)
    return sel32, imap


class Unpool(nn.Module):
    """x_fine[indices] = x_coarse, zeros elsewhere (orig :96-129)."""

    def forward(self, x_coarse: torch.Tensor, indices: torch.Tensor, num_nodes_fine: int) -> torch.Tensor:
        ops._require_cuda(x_coarse, indices)
        if x_coarse.dim() == 3:      # [batch, Nc, C] variant (orig :102, second branch)
            return torch.stack([self.forward(xc, indices, num_nodes_fine) for xc in x_coarse], dim=0)
        sel32, imap = _index_maps(indices, int(num_nodes_fine))
        return _b.UnpoolFn.apply(x_coarse, sel32, imap, int(num_nodes_fine))


class WeightedEdgeConv(nn.Module):
    """Message passing with a learned scalar weight per edge (orig :131-209)."""

    def __init__(self, in_dim: int, out_dim: int, aggr: str = "add"):
        super().__init__()
        self.in_dim, self.out_dim, self.aggr = in_dim, out_dim, aggr
        self.edge_weight_mlp = nn.Sequential(nn.Linear(2 * in_dim + 1, 64), nn.ReLU(), nn.Linear(64, 1), nn.Sigmoid())
        self.transform = nn.Linear(in_dim, out_dim)

    def _params(self):
        l0, l2 = self.edge_weight_mlp[0], self.edge_weight_mlp[2]
        return l0.weight, l0.bias, l2.weight, l2.bias, self.transform.weight, self.transform.bias

    def compute_edge_weights(self, x, edge_index, pos) -> torch.Tensor:
        """[E,1] weights sigmoid(MLP([x_src, x_dst, |pos_dst - pos_src|])) (orig :152)."""
        return _b.weighted_edge_conv(x, edge_index, pos, *self._params(), "add")[1]

    def forward(self, x, edge_index, pos, edge_weights=None, compute_weights: bool = True):
        """Returns (out [N,out_dim], edge_weights [E,1]) (orig :173)."""
        return _b.weighted_edge_conv(x, edge_index, pos, *self._params(), self.aggr, edge_weights=edge_weights,
                                     compute_weights=compute_weights)


class GMP(nn.Module):
    """One message-passing step with two-Linear MLPs and LayerNorm (orig :211-263):
        e' = e + edge_mlp([x_src, x_dst, e]);  x' = x + node_mlp([x, sum of e' onto receivers])
    Note the input order of edge_mlp: node rows first, edge row last (the opposite of EdgeBlock, mgnLayer.py:44)."""

    def __init__(self, node_dim: int, edge_dim: int, hidden_dim: int, activation: str = "relu"):
        super().__init__()
        self.dims = (node_dim, edge_dim, hidden_dim)
        self.activation = activation
        act = nn.ReLU if activation == "relu" else nn.SiLU
        self.edge_mlp = nn.Sequential(nn.Linear(2 * node_dim + edge_dim, hidden_dim), act(),
                                      nn.Linear(hidden_dim, edge_dim), nn.LayerNorm(edge_dim))
        self.node_mlp = nn.Sequential(nn.Linear(node_dim + edge_dim, hidden_dim), act(),
                                      nn.Linear(hidden_dim, node_dim), nn.LayerNorm(node_dim))

    def stack_config(self) -> StackConfig:
        if not (self.dims[0] == self.dims[1] == self.dims[2] == D):
            raise RuntimeError(f"the fused sm_100a path supports node_dim == edge_dim == hidden_dim == {D} only "
                               f"(got {self.dims})")
        # 'relu' runs on the fused block kernels; 'silu' (the bytecode's other choice, bistride_ops orig :216: ReLU if
        # activation == 'relu' else SiLU) has no derivative-from-output form and runs processor.eager_stack (announced)
        act = "relu" if self.activation == "relu" else "silu"
        return StackConfig(L_edge=0, L_node=0, act_edge=act, act_node=act, use_ln=True, mean=False)

    def step_weights(self, dtype: torch.dtype) -> StepWeights:
        def build():
            nd = self.dims[0]
            e0, e2, eln = self.edge_mlp[0], self.edge_mlp[2], self.edge_mlp[3]
            n0, n2, nln = self.node_mlp[0], self.node_mlp[2], self.node_mlp[3]
            w_s, w_d, w_e = e0.weight[:, :nd], e0.weight[:, nd:2 * nd], e0.weight[:, 2 * nd:]
            w_x, w_a = n0.weight[:, :nd], n0.weight[:, nd:]
            return ((w_e, [], e2.weight, e2.bias, eln.weight, eln.bias),
                    (w_a, [], n2.weight, n2.bias, nln.weight, nln.bias),
                    [w_s, w_d, w_x], [None, e0.bias, n0.bias])
        return cached_pack_step(self, dtype, build)

    def forward_csr(self, x, e_csr, plan: ops.GraphPlan):
        """The same step on edge rows already in the plan's receiver-CSR order; returns (x', e' in CSR order)."""
        return run_stack(self.stack_config(), plan, x, e_csr, [self.step_weights(x.dtype)])

    def forward(self, x, edge_attr, edge_index):
        ops._require_cuda(x, edge_attr, edge_index)
        plan = ops.PLAN_CACHE.get(edge_index, x.size(0))
        x, e_csr = self.forward_csr(x, permute_rows(edge_attr, plan.perm, plan.inv_perm), plan)
        return x, permute_rows(e_csr, plan.inv_perm, plan.perm)
