"""CPU oracle for the aero-gnn message-passing path.  TEST INFRASTRUCTURE ONLY.

This is a plain restatement (torch CPU tensors / numpy integers, no nn.Module, no autograd tricks) of the
reference algorithm.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import it; the product path (aero_gnn_b200) never does.

Pinning: the reference ships no tests or golden vectors.  The oracle is pinned by
tests/golden/*.pt, produced by oracle/gen_golden.py which imports the unmodified reference modules from
/root/reference (with torch_scatter / torch_geometric stand-ins, oracle/standins.py) and records their
outputs and autograd gradients; tests/test_oracle_golden.py checks every function below against them.
The torch_scatter / torch_geometric boundary itself is "parity unpinned" (packages absent, versions
un-pinned by the reference): their published semantics are restated in `scatter_add` / `scatter_mean`.

All functions take a `sd` state-dict-like mapping with the reference's parameter names and a key prefix.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ---- torch_scatter 2.x semantics (call sites mgnLayer.py:144,146; bsms_mgn.py:265-283) ----------
def scatter_add(src: torch.Tensor, index: torch.Tensor, dim_size: Optional[int] = None) -> torch.Tensor:
    n = int(index.max()) + 1 if dim_size is None else dim_size
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)      # CPU: sequential in index order == scatter_add_


def scatter_mean(src: torch.Tensor, index: torch.Tensor, dim_size: Optional[int] = None) -> torch.Tensor:
    s = scatter_add(src, index, dim_size)
    cnt = torch.zeros(s.shape[0], dtype=src.dtype, device=src.device).index_add_(
        0, index, torch.ones(index.numel(), dtype=src.dtype, device=src.device))
    return s / cnt.clamp(min=1).view(-1, *([1] * (src.dim() - 1)))


# ---- models/mlp.py:40-51 ------------------------------------------------------------------------
def mlp(sd: SD, prefix: str, x: torch.Tensor, act: str = "relu", use_ln: bool = True) -> torch.Tensor:
    n = 0
    while f"{prefix}layers.{n}.weight" in sd:
        n += 1
    f = getattr(F, act)
    for i in range(n):
        x = F.linear(x, sd[f"{prefix}layers.{i}.weight"], sd[f"{prefix}layers.{i}.bias"])
        if i < n - 1:
            x = f(x)
    if use_ln:
        w = sd[f"{prefix}layer_norm.weight"]
        x = F.layer_norm(x, (w.numel(),), w, sd[f"{prefix}layer_norm.bias"], 1e-5)
    return x


# ---- models/mgnLayer.py:32-49 (EdgeBlock) and :93-105 (EdgeBlockSum) ---------------------------
def edge_block(sd: SD, prefix: str, e, x, edge_index, act: str = "relu", use_ln: bool = True):
    row, col = edge_index[0].long(), edge_index[1].long()
    if f"{prefix}edge_lin" in sd:  # sum-trick form, always ReLU (mgnLayer.py:81)
        h = F.linear(e, sd[f"{prefix}edge_lin"]) + F.linear(x, sd[f"{prefix}src_lin"])[row] \
            + F.linear(x, sd[f"{prefix}dst_lin"], sd[f"{prefix}bias"])[col]
        # Sequential(ReLU, [Linear, ReLU] x L, Linear, LayerNorm?) -- mgnLayer.py:81-91
        ids = sorted(int(k[len(prefix) + 4:].split(".")[0]) for k in sd
                     if k.startswith(f"{prefix}mlp.") and k.endswith(".weight"))
        lin = [i for i in ids if sd[f"{prefix}mlp.{i}.weight"].dim() == 2]
        norm = [i for i in ids if sd[f"{prefix}mlp.{i}.weight"].dim() == 1]
        h = F.relu(h)
        for j, i in enumerate(lin):
            h = F.linear(h, sd[f"{prefix}mlp.{i}.weight"], sd[f"{prefix}mlp.{i}.bias"])
            if j < len(lin) - 1:
                h = F.relu(h)
        for i in norm:
            w = sd[f"{prefix}mlp.{i}.weight"]
            h = F.layer_norm(h, (w.numel(),), w, sd[f"{prefix}mlp.{i}.bias"], 1e-5)
        return h
    return mlp(sd, f"{prefix}mlp.", torch.cat([e, x[row], x[col]], dim=-1), act, use_ln)


# ---- models/mgnLayer.py:134-153 ------------------------------------------------------------------
def node_block(sd: SD, prefix: str, x, e, edge_index, aggregation: str = "add", act: str = "relu", use_ln: bool = True):
    col = edge_index[1].long()
    if aggregation == "mean":
        agg = scatter_mean(e, col, x.size(0))
    elif aggregation == "add":
        agg = scatter_add(e, col, x.size(0))
    else:
        raise ValueError(f"Unsupported aggregation method: {aggregation}")
    return mlp(sd, f"{prefix}mlp.", torch.cat([x, agg], dim=-1), act, use_ln)


# ---- models/mgnLayer.py:177-213 ------------------------------------------------------------------
def mgn_layer(sd: SD, prefix: str, x, e, edge_index, aggregation="add", act="relu"):
    e = e + edge_block(sd, f"{prefix}edge_block.", e, x, edge_index, act)
    x = x + node_block(sd, f"{prefix}node_block.", x, e, edge_index, aggregation, act)
    return x, e


def _count(sd: SD, fmt: str) -> int:
    n = 0
    while any(k.startswith(fmt.format(n)) for k in sd):
        n += 1
    return n


# ---- models/mgn.py:108-139 -----------------------------------------------------------------------
def mgn_forward(sd: SD, node_attr, edge_attr, edge_index, aggregation="add", act="relu"):
    x = mlp(sd, "node_encoder.", node_attr, act)
    e = mlp(sd, "edge_encoder.", edge_attr, act)
    for i in range(_count(sd, "layers.{}.")):
        x, e = mgn_layer(sd, f"layers.{i}.", x, e, edge_index, aggregation, act)
    return mlp(sd, "decoder.", x, act, use_ln=False)


# ---- models/fouriermgn.py:111-183 ----------------------------------------------------------------
def fourier_embedding(pos, dim=2, start=-3, length=7):
    xs = pos[:, :dim]
    k = torch.arange(start, start + length, dtype=pos.dtype, device=pos.device)
    ph = ((2.0 ** k) * math.pi).view(1, 1, -1) * xs.unsqueeze(-1)
    return torch.cat([torch.cos(ph), torch.sin(ph)], dim=-1).reshape(pos.shape[0], -1)


def fourier_mgn_forward(sd: SD, node_attr, edge_attr, edge_index, aggregation="add", act="relu", dim=2, start=-3,
                        length=7):
    xin = torch.cat([node_attr, fourier_embedding(node_attr, dim, start, length)], dim=-1)
    return mgn_forward(sd, xin, edge_attr, edge_index, aggregation, act)


# ---- models/poolmgn.py:115-157 --------------------------------------------------------------------
def pool_mgn_forward(sd: SD, node_attr, edge_attr, edge_index, batch=None, method="mean", aggregation="add", act="relu"):
    g = mlp(sd, "global_encoder.", node_attr, act, use_ln=False)
    if batch is None:
        batch = torch.zeros(node_attr.size(0), dtype=torch.long, device=node_attr.device)
    nb = int(batch.max()) + 1
    if method == "mean":
        pooled = scatter_mean(g, batch, nb)
    elif method == "add":
        pooled = scatter_add(g, batch, nb)
    else:
        pooled = torch.stack([g[batch == b].max(dim=0).values for b in range(nb)])
    xin = torch.cat((node_attr, pooled[batch]), dim=-1)
    return mgn_forward(sd, xin, edge_attr, edge_index, aggregation, act)


# ---- models/bsms_mgn.py:217-301 (integer part in numpy, bit-exact contract) -----------------------
def stride_pool_indices(batch: np.ndarray, posx: Optional[np.ndarray], stride: int):
    """fine_to_coarse, coarse_batch.  Per graph (consecutive ids): stable argsort of pos[:,0], rank // stride."""
    batch = np.asarray(batch, dtype=np.int64)
    n = batch.shape[0]
    f2c = np.empty(n, dtype=np.int64)
    cb = []
    off = 0
    if n:
        starts = np.flatnonzero(np.concatenate([[True], batch[1:] != batch[:-1]]))
        for s, t in zip(starts, list(starts[1:]) + [n]):
            idx = np.arange(s, t)
            if posx is not None:
                idx = idx[np.argsort(np.asarray(posx)[s:t], kind="stable")]
            local = np.arange(t - s) // stride
            nc = int(local[-1]) + 1
            f2c[idx] = local + off
            cb.append(np.full(nc, batch[s], dtype=np.int64))
            off += nc
    return f2c, (np.concatenate(cb) if cb else np.empty(0, dtype=np.int64))


def coarsen_edge_indices(edge_index: np.ndarray, f2c: np.ndarray, nc: int):
    """coarse_edge_index (sorted unique (sender, receiver) pairs, self-loops kept), inverse."""
    m = max(nc, 1)
    key = f2c[edge_index[0]] * m + f2c[edge_index[1]]
    uniq, inv = np.unique(key, return_inverse=True)
    return np.stack([uniq // m, uniq % m]).astype(np.int64), inv.astype(np.int64)


def downsample(x, e, edge_index, batch, pos, stride: int):
    # integer part on the host in numpy (the bit-exact contract); tensors may live on any device (the tests run the
    # same restatement in bf16 on the GPU as "the reference's own bf16 mode", train.py:30-33)
    dev = x.device
    f2c_np, cb_np = stride_pool_indices(batch.cpu().numpy(),
                                        None if pos is None else pos[:, 0].float().cpu().numpy(), stride)
    nc = cb_np.shape[0]
    f2c = torch.from_numpy(f2c_np).to(dev)
    cx = scatter_mean(x, f2c, nc)
    cpos = scatter_mean(pos, f2c, nc) if pos is not None else None
    cei_np, inv_np = coarsen_edge_indices(edge_index.cpu().numpy(), f2c_np, nc)
    if cei_np.shape[1] > 0:
        ce = scatter_mean(e, torch.from_numpy(inv_np).to(dev))
    else:
        ce = e.new_zeros((0, e.size(1)))
    return cx, ce, torch.from_numpy(cei_np).to(dev), torch.from_numpy(cb_np).to(dev), cpos, f2c


# ---- models/bsms_mgn.py:126-215 --------------------------------------------------------------------
def bsms_forward(sd: SD, node_attr, edge_attr, edge_index, batch=None, pos=None, stride=2, aggregation="add", act="relu"):
    if batch is None:
        batch = torch.zeros(node_attr.size(0), dtype=torch.long, device=node_attr.device)
    x = mlp(sd, "node_encoder.", node_attr, act)
    e = mlp(sd, "edge_encoder.", edge_attr, act)
    ei, b, p = edge_index, batch, pos
    skips, assigns = [], []
    for s in range(_count(sd, "down_layers.{}.")):
        for k in range(_count(sd, f"down_layers.{s}." + "{}.")):
            x, e = mgn_layer(sd, f"down_layers.{s}.{k}.", x, e, ei, aggregation, act)
        skips.append((x, e, ei, b, p))
        x, e, ei, b, p, a = downsample(x, e, ei, b, p, stride)
        assigns.append(a)
    for k in range(_count(sd, "bottleneck_layers.{}.")):
        x, e = mgn_layer(sd, f"bottleneck_layers.{k}.", x, e, ei, aggregation, act)
    for s in range(_count(sd, "up_layers.{}.")):
        sx, se, sei, sb, sp = skips[-(s + 1)]
        x = x[assigns[-(s + 1)]] + sx
        e, ei, b, p = se, sei, sb, sp
        for k in range(_count(sd, f"up_layers.{s}." + "{}.")):
            x, e = mgn_layer(sd, f"up_layers.{s}.{k}.", x, e, ei, aggregation, act)
    return mlp(sd, "decoder.", x, act, use_ln=False)


# ---- graph plan (new component: checked against torch.sort(stable=True)) ---------------------------
def receiver_csr(edge_index: np.ndarray, n: int):
    dst = edge_index[1]
    perm = np.argsort(dst, kind="stable")
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, dst + 1, 1)
    rowptr = np.cumsum(rowptr)
    src_csr = edge_index[0][perm]
    sperm = np.argsort(src_csr, kind="stable")
    sptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(sptr, src_csr + 1, 1)
    return rowptr, perm, src_csr, dst[perm], np.cumsum(sptr), sperm
