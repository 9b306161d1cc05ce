"""In-process stand-ins for torch_scatter / torch_geometric so the UNMODIFIED reference modules import.
TEST INFRASTRUCTURE ONLY (used by gen_golden.py and bench.py --impl reference; needs the reference
checkout on sys.path, i.e. only where /root/reference exists).

Semantics restated from torch_scatter 2.x: zeros-init, scatter_add_ along dim 0, mean = sum / count.clamp(1),
dim_size=None -> index.max()+1; torch_geometric global_*_pool = segment reduce by `batch`.
"""
import sys
import types

import torch


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0
    n = int(index.max()) + 1 if dim_size is None else dim_size
    res = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return res.scatter_add_(0, idx, src)


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    s = scatter_add(src, index, dim, None, dim_size)
    cnt = scatter_add(torch.ones(src.shape[0], dtype=src.dtype, device=src.device), index, 0, None, s.shape[0])
    return s / cnt.clamp(min=1).view(-1, *([1] * (src.dim() - 1)))


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    return scatter_mean(src, index, dim, out, dim_size) if reduce == "mean" else scatter_add(src, index, dim, out, dim_size)


def global_add_pool(x, batch, size=None):
    return scatter_add(x, batch, 0, None, size)


def global_mean_pool(x, batch, size=None):
    return scatter_mean(x, batch, 0, None, size)


def global_max_pool(x, batch, size=None):
    n = int(batch.max()) + 1 if size is None else size
    return torch.stack([x[batch == b].max(dim=0).values for b in range(n)])


def install(reference_root="/root/reference"):
    ts = types.ModuleType("torch_scatter")
    ts.scatter_add, ts.scatter_mean, ts.scatter = scatter_add, scatter_mean, scatter
    tg = types.ModuleType("torch_geometric")
    tgn = types.ModuleType("torch_geometric.nn")
    tgn.global_add_pool, tgn.global_mean_pool, tgn.global_max_pool = global_add_pool, global_mean_pool, global_max_pool
    for name in ("MessagePassing", "GCNConv", "GINEConv", "GraphSAGE"):
        setattr(tgn, name, type(name, (torch.nn.Module,), {}))
    tg.nn = tgn
    sys.modules.setdefault("torch_scatter", ts)
    sys.modules.setdefault("torch_geometric", tg)
    sys.modules.setdefault("torch_geometric.nn", tgn)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    # the reference's `models` directory has no __init__.py (namespace package) and this repository ships a regular
    # `models` shim package, which would win the import: pin `models` to the reference directory explicitly
    ref_models = types.ModuleType("models")
    ref_models.__path__ = [reference_root + "/models"]
    sys.modules["models"] = ref_models
