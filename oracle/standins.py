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


# ---- torch_geometric data plumbing + pyvista / matplotlib, only for running the reference's train.py unchanged -------
class Data:
    """Attribute bag with the torch_geometric.data.Data surface train.py / utils.py touch."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def to(self, device, *a, **k):
        out = type(self)()
        for key, v in self.__dict__.items():
            setattr(out, key, v.to(device, *a, **k) if torch.is_tensor(v) else v)
        return out

    @property
    def num_nodes(self):
        return self.x.shape[0]


class Dataset(torch.utils.data.Dataset):
    pass


def collate(items):
    """PyG disjoint-union batching (train.py:50-51): node rows / edge rows concatenated, edge_index offset per graph,
    `batch` = graph id of every node."""
    out, off, batch = Data(), 0, []
    keys = items[0].keys()
    for k in keys:
        vals = [getattr(d, k) for d in items]
        if k == "edge_index":
            offs = [0]
            for d in items[:-1]:
                offs.append(offs[-1] + d.num_nodes)
            setattr(out, k, torch.cat([v + o for v, o in zip(vals, offs)], dim=1))
        elif torch.is_tensor(vals[0]) and vals[0].dim() >= 1 and vals[0].shape[0] in (items[0].num_nodes, items[0].edge_index.shape[1]):
            setattr(out, k, torch.cat(vals, dim=0))
        else:
            setattr(out, k, vals)
    for i, d in enumerate(items):
        batch.append(torch.full((d.num_nodes,), i, dtype=torch.long))
        off += d.num_nodes
    out.batch = torch.cat(batch)
    out.num_graphs = len(items)
    return out


class DataLoader:
    def __init__(self, dataset, batch_size=1, shuffle=False, **_):
        self.dataset, self.batch_size, self.shuffle = list(dataset), int(batch_size), shuffle

    def __len__(self):
        return -(-len(self.dataset) // self.batch_size)

    def __iter__(self):
        idx = torch.randperm(len(self.dataset)).tolist() if self.shuffle else list(range(len(self.dataset)))
        for i in range(0, len(idx), self.batch_size):
            yield collate([self.dataset[j] for j in idx[i: i + self.batch_size]])


def to_undirected(edge_index, *a, **k):
    both = torch.cat([edge_index, edge_index.flip(0)], dim=1)
    n = int(both.max()) + 1 if both.numel() else 1
    key = torch.unique(both[0] * n + both[1])
    return torch.stack([key // n, key % n])


def is_undirected(edge_index, *a, **k):
    n = int(edge_index.max()) + 1 if edge_index.numel() else 1
    a_ = torch.sort(edge_index[0] * n + edge_index[1]).values
    b_ = torch.sort(edge_index[1] * n + edge_index[0]).values
    return bool(torch.equal(a_, b_))


class _Noop(types.ModuleType):
    """A module whose every attribute is a callable returning another no-op (pyvista, matplotlib.pyplot)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _NoopCallable()


class _NoopCallable:
    def __call__(self, *a, **k):
        return _NoopCallable()

    def __getattr__(self, name):
        return _NoopCallable()


def install_full():
    """Everything train.py / utils.py / dataset.py / inference.py import that this image lacks, on top of install()."""
    tg = sys.modules["torch_geometric"]
    for name, members in {"torch_geometric.loader": dict(DataLoader=DataLoader),
                          "torch_geometric.data": dict(Data=Data, Dataset=Dataset, Batch=Data),
                          "torch_geometric.utils": dict(to_undirected=to_undirected, is_undirected=is_undirected)}.items():
        m = types.ModuleType(name)
        for k, v in members.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
        setattr(tg, name.split(".")[-1], sys.modules[name])
    sys.modules.setdefault("pyvista", _Noop("pyvista"))
    mpl = sys.modules.setdefault("matplotlib", _Noop("matplotlib"))
    sys.modules.setdefault("matplotlib.pyplot", _Noop("matplotlib.pyplot"))
    mpl.pyplot = sys.modules["matplotlib.pyplot"]


def install(reference_root="/root/reference", pin_models=True):
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
    # (pin_models=False leaves `models` to the repository's shim: the reference's train.py on OUR models)
    if pin_models:
        ref_models = types.ModuleType("models")
        ref_models.__path__ = [reference_root + "/models"]
        sys.modules["models"] = ref_models
