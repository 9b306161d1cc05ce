"""CPU oracle for the BFS-bistride operators.  TEST INFRASTRUCTURE ONLY (same rules as mgn_oracle.py).

Upstream ships these components only as CPython 3.11 bytecode
(/root/reference/models/__pycache__/bistride_ops.cpython-311.pyc and the stale bsms_mgn.cpython-311.pyc): no source, no
test, no golden vector, and the bytecode does not run on this image's Python 3.12.  The functions below restate the
behaviour decoded from the marshal stream (oracle/decode_bistride_pyc.py prints -- and
oracle/bistride_pyc_decoded.txt holds -- the constants, names and operand order: hidden width 64, fallback ratio 0.3,
the cat orders [x_src, x_dst, len] / [x_src, x_dst, e] / [x, agg], pos[dst] - pos[src], `a = a + b` residuals).

PINNING: oracle/pyc311_vm.py is a small interpreter for 3.11 bytecode; oracle/gen_bistride_golden.py uses it to
EXECUTE the reference's own code objects (real torch / nn.Module objects, torch_scatter stand-in, the reference's real
models/mlp.py) and records BFS distances, selections, hierarchies, Unpool, WeightedEdgeConv, GMP and BSMS_MeshGraphNet
outputs with autograd gradients into tests/golden/bistride.pt.  tests/test_bistride_golden.py checks every function
below against those vectors (integers exact, floats <= 1e-5, gradients <= 1e-4).  What stays unpinned is only the
torch_scatter boundary (absent package, published semantics restated -- as for mgn_oracle.py).
"orig :NN" = first line of the code object in the pyc.  Plain torch CPU tensors, Python loops for the BFS.
"""
from __future__ import annotations

from collections import deque
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .mgn_oracle import scatter_add, scatter_mean, mlp

SD = Dict[str, torch.Tensor]


# ---- BistridePooling.bfs_distance (orig :21) ----------------------------------------------------
def bfs_distance(edge_index: torch.Tensor, num_nodes: int, start_node: int) -> torch.Tensor:
    dist = torch.full((num_nodes,), -1, dtype=torch.long)
    dist[start_node] = 0
    adj = [[] for _ in range(num_nodes)]
    src, dst = edge_index[0].tolist(), edge_index[1].tolist()
    for s, d in zip(src, dst):
        adj[s].append(d)
    q = deque([start_node])
    d = dist.tolist()
    while q:
        u = q.popleft()
        for v in adj[u]:
            if d[v] == -1:
                d[v] = d[u] + 1
                q.append(v)
    return torch.tensor(d, dtype=torch.long)


# ---- BistridePooling.select_bistride_nodes (orig :56) -------------------------------------------
def select_bistride_nodes(edge_index: torch.Tensor, num_nodes: int, pos: Optional[torch.Tensor] = None) -> torch.Tensor:
    if pos is not None:
        center = pos.mean(dim=0)
        seed = int(torch.argmin(torch.norm(pos - center, dim=1)))
    else:
        seed = int(torch.argmax(torch.bincount(edge_index[0], minlength=num_nodes)))
    d = bfs_distance(edge_index, num_nodes, seed)
    sel = torch.where((d % 2 == 0) & (d >= 0))[0]
    if len(sel) < num_nodes * 0.3:
        sel = torch.where(d >= 0)[0]
    return sel


# ---- Unpool.forward (orig :102) -----------------------------------------------------------------
def unpool(x_coarse: torch.Tensor, indices: torch.Tensor, num_nodes_fine: int) -> torch.Tensor:
    if x_coarse.dim() == 2:
        out = torch.zeros(num_nodes_fine, x_coarse.shape[1], dtype=x_coarse.dtype)
        out[indices] = x_coarse
        return out
    out = torch.zeros(x_coarse.shape[0], num_nodes_fine, x_coarse.shape[2], dtype=x_coarse.dtype)
    out[:, indices, :] = x_coarse
    return out


# ---- WeightedEdgeConv (orig :131-209) -----------------------------------------------------------
def wec_edge_weights(sd: SD, prefix: str, x, edge_index, pos) -> torch.Tensor:
    src, dst = edge_index[0].long(), edge_index[1].long()
    length = torch.norm(pos[dst] - pos[src], dim=1, keepdim=True)
    feat = torch.cat([x[src], x[dst], length.to(x.dtype)], dim=1)
    h = F.relu(F.linear(feat, sd[f"{prefix}edge_weight_mlp.0.weight"], sd[f"{prefix}edge_weight_mlp.0.bias"]))
    return torch.sigmoid(F.linear(h, sd[f"{prefix}edge_weight_mlp.2.weight"], sd[f"{prefix}edge_weight_mlp.2.bias"]))


def wec(sd: SD, prefix: str, x, edge_index, pos, edge_weights=None, compute_weights=True, aggr="add"):
    src, dst = edge_index[0].long(), edge_index[1].long()
    if compute_weights and edge_weights is None:
        edge_weights = wec_edge_weights(sd, prefix, x, edge_index, pos)
    xt = F.linear(x, sd[f"{prefix}transform.weight"], sd[f"{prefix}transform.bias"])
    msg = xt[src] * edge_weights
    if aggr == "add":
        out = scatter_add(msg, dst, x.shape[0])
    elif aggr == "mean":
        out = scatter_mean(msg, dst, x.shape[0])
    else:
        raise ValueError(f"Unknown aggregation: {aggr}")
    return out, edge_weights


# ---- GMP (orig :211-263) ------------------------------------------------------------------------
def _seq_mlp(sd: SD, prefix: str, x, act):
    h = act(F.linear(x, sd[f"{prefix}0.weight"], sd[f"{prefix}0.bias"]))
    y = F.linear(h, sd[f"{prefix}2.weight"], sd[f"{prefix}2.bias"])
    w = sd[f"{prefix}3.weight"]
    return F.layer_norm(y, (w.numel(),), w, sd[f"{prefix}3.bias"], 1e-5)


def gmp(sd: SD, prefix: str, x, edge_attr, edge_index, activation="relu"):
    act = F.relu if activation == "relu" else F.silu
    src, dst = edge_index[0].long(), edge_index[1].long()
    e_in = torch.cat([x[src], x[dst], edge_attr], dim=1)
    edge_attr = edge_attr + _seq_mlp(sd, f"{prefix}edge_mlp.", e_in, act)
    agg = scatter_add(edge_attr, dst, x.shape[0])
    x = x + _seq_mlp(sd, f"{prefix}node_mlp.", torch.cat([x, agg], dim=1), act)
    return x, edge_attr


# ---- MultiScaleGraphPreprocessor.create_multiscale_graph (bsms_mgn pyc orig :32) ----------------
def create_multiscale_graph(edge_index: torch.Tensor, pos: torch.Tensor, num_levels: int = 3):
    n = pos.shape[0]
    multi = {"edge_indices": [edge_index], "node_indices": [], "num_nodes": [n], "positions": [pos]}
    for _ in range(num_levels):
        sel = select_bistride_nodes(edge_index, n, pos)
        multi["node_indices"].append(sel)
        index_map = torch.full((n,), -1, dtype=torch.long)
        index_map[sel] = torch.arange(len(sel))
        pos = pos[sel]
        multi["positions"].append(pos)
        src, dst = edge_index[0], edge_index[1]
        mask = (index_map[src] >= 0) & (index_map[dst] >= 0)
        ns, nd = index_map[src[mask]], index_map[dst[mask]]
        mask = ns != nd
        edge_index = torch.stack([ns[mask], nd[mask]], dim=0)
        multi["edge_indices"].append(edge_index)
        n = len(sel)
        multi["num_nodes"].append(n)
    return multi


# ---- BSMSGMP.forward (bsms_mgn pyc orig :145) ---------------------------------------------------
def bsmsgmp(sd: SD, prefix: str, num_levels: int, x, edge_attrs, edge_indices, node_indices, num_nodes_list, positions):
    edge_attrs = list(edge_attrs)
    skips, w_down = [], []
    for i in range(num_levels):
        x, edge_attrs[i] = gmp(sd, f"{prefix}down_gmps.{i}.", x, edge_attrs[i], edge_indices[i])
        skips.append(x.clone())
        xc, ew = wec(sd, f"{prefix}down_edge_convs.{i}.", x, edge_indices[i], positions[i], compute_weights=True)
        w_down.append(ew)
        x = x + xc
        x = x[node_indices[i]]
    x, edge_attrs[-1] = gmp(sd, f"{prefix}bottom_gmp.", x, edge_attrs[-1], edge_indices[-1])
    for i in range(num_levels - 1, -1, -1):
        x = unpool(x, node_indices[i], num_nodes_list[i])
        xc, _ = wec(sd, f"{prefix}up_edge_convs.{i}.", x, edge_indices[i], positions[i], edge_weights=w_down[i],
                    compute_weights=False)
        x = x + xc + skips[i]
    return x


# ---- BSMS_MeshGraphNet.forward (bsms_mgn pyc orig :281) -----------------------------------------
def bsms_meshgraphnet(sd: SD, num_levels: int, node_attr, edge_attr, multi, act: str = "relu"):
    x = mlp(sd, "node_encoder.", node_attr, act)
    e = mlp(sd, "edge_encoder.", edge_attr, act)
    edge_attrs = [e] + [torch.zeros(ei.shape[1], x.shape[1], dtype=x.dtype) for ei in multi["edge_indices"][1:]]
    x = bsmsgmp(sd, "bsgmp.", num_levels, x, edge_attrs, multi["edge_indices"], multi["node_indices"],
                multi["num_nodes"], multi["positions"])
    return mlp(sd, "decoder.", x, act, use_ln=False)
