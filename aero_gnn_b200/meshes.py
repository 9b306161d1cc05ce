"""Synthetic meshes with the input layout of the reference's dataset (dataset.py:39-106, utils.py:39-40).

There is no network and no dataset on the GPU box, so bench.py and the tests build meshes here:
  * airfoil_o_mesh  -- 2-D structured O-mesh around a NACA-0012 section, quads split by one diagonal
                       (N = n_theta*n_r, E = 2*n_theta*(3*n_r - 2) directed edges)
  * wing_surface_mesh -- 3-D tapered wing surface, periodic around the section, open along the span
                       (N = nu*nv, E = 2*nu*(3*nv - 2))
Directed edge lists hold both directions, coalesced and sorted by (sender, receiver) like
torch_geometric.utils.to_undirected.  Node features [pos, normal, mach, alpha] (2-D) or [pos, normal]
(3-D); edge features [target - source, |d|]; every column z-scored (dataset.py:403-409).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch


@dataclass
class Mesh:
    pos: torch.Tensor          # [N, p] float32
    edge_index: torch.Tensor   # [2, E] int64, sorted by (sender, receiver)
    node_attr: torch.Tensor    # [N, F] float32, z-scored
    edge_attr: torch.Tensor    # [E, p+1] float32, z-scored
    batch: torch.Tensor        # [N] int64
    target: Optional[torch.Tensor] = None

    @property
    def num_nodes(self) -> int:
        return int(self.pos.shape[0])

    @property
    def num_edges(self) -> int:
        return int(self.edge_index.shape[1])

    def to(self, device, dtype=None) -> "Mesh":
        f = lambda t: None if t is None else (t.to(device=device, dtype=dtype) if (dtype and t.is_floating_point()) else t.to(device))
        return Mesh(f(self.pos), f(self.edge_index), f(self.node_attr), f(self.edge_attr), f(self.batch), f(self.target))


def _zscore(a: np.ndarray) -> np.ndarray:
    mu = a.mean(axis=0, keepdims=True)
    sd = a.std(axis=0, keepdims=True)
    sd = np.where(sd < 1e-12, 1.0, sd)
    return ((a - mu) / sd).astype(np.float32)


def _structured_edges(ni: int, nj: int, periodic_i: bool = True) -> np.ndarray:
    """Directed edges of an ni x nj grid (i periodic), each quad split by the (i,j)-(i+1,j+1) diagonal."""
    i, j = np.meshgrid(np.arange(ni), np.arange(nj), indexing="ij")
    idx = lambda a, b: (a % ni) * nj + b
    pairs = []
    # along i
    if periodic_i:
        pairs.append((idx(i, j).ravel(), idx(i + 1, j).ravel()))
    else:
        m = i < ni - 1
        pairs.append((idx(i, j)[m], idx(i + 1, j)[m]))
    # along j
    m = j < nj - 1
    pairs.append((idx(i, j)[m], idx(i, j + 1)[m]))
    # diagonal
    if periodic_i:
        pairs.append((idx(i, j)[m], idx(i + 1, j + 1)[m]))
    else:
        m2 = m & (i < ni - 1)
        pairs.append((idx(i, j)[m2], idx(i + 1, j + 1)[m2]))
    a = np.concatenate([p[0] for p in pairs])
    b = np.concatenate([p[1] for p in pairs])
    s = np.concatenate([a, b]).astype(np.int64)
    r = np.concatenate([b, a]).astype(np.int64)
    n = ni * nj
    key = np.unique(s * n + r)            # coalesce + sort by (sender, receiver)
    return np.stack([key // n, key % n])


def _naca0012(theta: np.ndarray):
    """Closed NACA-0012 contour parametrised by angle; returns x, y, outward normal."""
    x = 0.5 * (1.0 + np.cos(theta))
    yt = 0.6 * (0.2969 * np.sqrt(np.maximum(x, 0)) - 0.1260 * x - 0.3516 * x ** 2 + 0.2843 * x ** 3 - 0.1036 * x ** 4)
    y = np.where(np.sin(theta) >= 0, yt, -yt)
    dx = np.gradient(x)
    dy = np.gradient(y)
    nrm = np.stack([dy, -dx], axis=-1)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=-1, keepdims=True), 1e-12)
    return x, y, nrm


def airfoil_o_mesh(n_theta: int = 100, n_r: int = 50, seed: int = 0, mach: float = 0.3, alpha: float = 2.0,
                   out_dim: int = 4) -> Mesh:
    rng = np.random.default_rng(seed)
    theta = np.linspace(0.0, 2.0 * np.pi, n_theta, endpoint=False) + 1e-3
    xs, ys, nrm = _naca0012(theta)
    stretch = 1.08 + 0.02 * rng.random()
    rad = (stretch ** np.arange(n_r) - 1.0) / (stretch - 1.0) * 0.02        # geometric wall-normal spacing
    cx, cy = 0.5, 0.0
    dirx, diry = xs - cx, ys - cy
    nd = np.maximum(np.hypot(dirx, diry), 1e-9)
    px = xs[:, None] + (dirx / nd)[:, None] * rad[None, :]
    py = ys[:, None] + (diry / nd)[:, None] * rad[None, :]
    pos = np.stack([px.ravel(), py.ravel()], axis=-1)
    # deterministic jitter << cell size so pos[:,0] has no ties (bistride sort is then well defined)
    n = pos.shape[0]
    pos[:, 0] += (np.arange(n) * 0.6180339887498949 % 1.0) * 1e-6
    pos = pos.astype(np.float32)
    normal = np.repeat(nrm[:, None, :], n_r, axis=1).reshape(-1, 2)
    ei = _structured_edges(n_theta, n_r, periodic_i=True)
    feats = np.concatenate([pos, normal, np.full((n, 1), mach), np.full((n, 1), alpha)], axis=1)
    feats[:, 4] += 1e-3 * rng.standard_normal(n)   # keep constant columns from being degenerate after z-scoring
    feats[:, 5] += 1e-3 * rng.standard_normal(n)
    d = pos[ei[1]] - pos[ei[0]]
    eattr = np.concatenate([d, np.linalg.norm(d, axis=1, keepdims=True)], axis=1)
    tgt = np.stack([np.sin(3 * pos[:, 0]) * np.cos(2 * pos[:, 1]), pos[:, 0] * pos[:, 1], np.cos(5 * pos[:, 1]),
                    np.tanh(pos[:, 0])][:out_dim], axis=-1)
    return Mesh(torch.from_numpy(pos), torch.from_numpy(ei), torch.from_numpy(_zscore(feats)),
                torch.from_numpy(_zscore(eattr)), torch.zeros(n, dtype=torch.long), torch.from_numpy(_zscore(tgt)))


def wing_surface_mesh(nu: int = 1000, nv: int = 1000, out_dim: int = 5) -> Mesh:
    """3-D tapered wing surface: nu points around the section (periodic) x nv span stations (open)."""
    theta = np.linspace(0.0, 2.0 * np.pi, nu, endpoint=False) + 1e-3
    xs, ys, nrm2 = _naca0012(theta)
    span = np.linspace(0.0, 4.0, nv)
    chord = 1.0 - 0.15 * span                                   # linear taper
    px = xs[:, None] * chord[None, :] + 0.25 * span[None, :]    # sweep
    py = ys[:, None] * chord[None, :]
    pz = np.broadcast_to(span[None, :], px.shape)
    pos = np.stack([px.ravel(), py.ravel(), pz.ravel()], axis=-1)
    n = pos.shape[0]
    pos[:, 0] += (np.arange(n) * 0.6180339887498949 % 1.0) * 1e-6
    pos = pos.astype(np.float32)
    normal = np.concatenate([np.repeat(nrm2[:, None, :], nv, axis=1).reshape(-1, 2), np.zeros((n, 1))], axis=1)
    ei = _structured_edges(nu, nv, periodic_i=True)
    feats = np.concatenate([pos, normal], axis=1)
    feats[:, 5] += 1e-3 * np.sin(np.arange(n) * 0.37)
    d = pos[ei[1]] - pos[ei[0]]
    eattr = np.concatenate([d, np.linalg.norm(d, axis=1, keepdims=True)], axis=1)
    tgt = np.stack([np.sin(3 * pos[:, 0]), np.cos(2 * pos[:, 1]), pos[:, 2] * 0.1, np.tanh(pos[:, 0]),
                    np.sin(pos[:, 2])][:out_dim], axis=-1)
    return Mesh(torch.from_numpy(pos), torch.from_numpy(ei), torch.from_numpy(_zscore(feats)),
                torch.from_numpy(_zscore(eattr)), torch.zeros(n, dtype=torch.long), torch.from_numpy(_zscore(tgt)))


def batch_meshes(meshes: List[Mesh]) -> Mesh:
    """Disjoint-union batching (torch_geometric DataLoader collate, train.py:50-51)."""
    off = 0
    pos, ei, na, ea, bt, tg = [], [], [], [], [], []
    for g, m in enumerate(meshes):
        pos.append(m.pos); na.append(m.node_attr); ea.append(m.edge_attr)
        ei.append(m.edge_index + off)
        bt.append(torch.full((m.num_nodes,), g, dtype=torch.long))
        if m.target is not None:
            tg.append(m.target)
        off += m.num_nodes
    return Mesh(torch.cat(pos), torch.cat(ei, dim=1), torch.cat(na), torch.cat(ea), torch.cat(bt),
                torch.cat(tg) if tg else None)


def random_graph(n: int, e: int, seed: int = 0, width: int = 128, dtype=torch.float32):
    """Unstructured random multigraph (duplicates and self-loops allowed) for kernel-level tests."""
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, max(n, 1), (2, e), generator=g, dtype=torch.long)
    x = torch.randn(n, width, generator=g).to(dtype)
    ea = torch.randn(e, width, generator=g).to(dtype)
    return x, ea, ei
