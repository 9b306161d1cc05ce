"""Golden vectors for the BFS-bistride components, produced by EXECUTING THE REFERENCE'S OWN BYTECODE.
TEST INFRASTRUCTURE ONLY; runs only where /root/reference exists (this container), writes tests/golden/bistride.pt.

    python oracle/gen_bistride_golden.py

Upstream ships `models.bistride_ops` and the older `models.bsms_mgn` only as CPython 3.11 bytecode.  oracle/pyc311_vm.py
interprets that bytecode (torch / nn.Module objects are the real ones of this interpreter; torch_scatter is the
stand-in of oracle/standins.py, models.mlp is the reference's real source).  Everything recorded here is therefore the
reference's behaviour, not a restatement: BFS distances, selected nodes, the multi-level hierarchy, Unpool,
WeightedEdgeConv (computed and reused weights, add / mean), GMP, BSMSGMP through BSMS_MeshGraphNet, with autograd
gradients.  tests/test_oracle_golden.py pins oracle/bistride_oracle.py to these vectors; tests/test_gpu_bistride.py
compares the CUDA path with them.
"""
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

import numpy as np   # noqa: E402
import torch         # noqa: E402

from oracle import pyc311_vm as V   # noqa: E402
from oracle import standins         # noqa: E402


def load_reference_modules():
    standins.install(REF)
    pyc = REF + "/models/__pycache__/"
    g1 = V.exec_module(pyc + "bistride_ops.cpython-311.pyc", "models.bistride_ops")
    m1 = types.ModuleType("models.bistride_ops")
    m1.__dict__.update(g1)

    def importer(name, fromlist, level):
        if name == "models.bistride_ops":
            return m1
        mod = importlib.import_module(name)
        return mod if fromlist else sys.modules[name.split(".")[0]]

    g2 = V.exec_module(pyc + "bsms_mgn.cpython-311.pyc", "models.bsms_mgn", importer)
    return g1, g2


def checksums(state):
    """Per-parameter (sum, norm) in float64: pins a state_dict without storing it.  The mirror modules of
    aero_gnn_b200.models draw their initial parameters in the same order as the reference, so the tests rebuild the
    state from the recorded seed and check it against these numbers."""
    return {k: (float(v.double().sum()), float(v.double().norm())) for k, v in state.items()}


def main():
    from aero_gnn_b200.meshes import airfoil_o_mesh
    B, S = load_reference_modules()
    out = {}
    mesh = airfoil_o_mesh(24, 12, seed=0)      # 288 nodes, 1,656 edges: keeps the fixture small
    n, ei = mesh.num_nodes, mesh.edge_index
    gen = torch.Generator().manual_seed(7)
    pos = mesh.pos[:, :2].clone() + 1e-4 * torch.rand(n, 2, generator=gen)
    out["mesh"] = dict(edge_index=ei, pos=pos, node_attr=mesh.node_attr, edge_attr=mesh.edge_attr, n=n)

    # ---- BistridePooling ----
    rng = np.random.default_rng(3)
    rnd = torch.from_numpy(rng.integers(0, 300, size=(2, 700)).astype(np.int64))     # directed, duplicates, self-loops
    out["bfs"] = [dict(edge_index=e, n=m, start=s, dist=B["BistridePooling"].bfs_distance(e, m, s))
                  for e, m, s in ((ei, n, 0), (ei, n, 217), (rnd, 300, 5), (rnd, 300, 299))]
    star = torch.tensor([(0, k) for k in range(1, 10)] + [(k, 0) for k in range(1, 10)]).t().contiguous()
    out["select"] = [dict(edge_index=e, n=m, pos=p, selected=B["BistridePooling"].select_bistride_nodes(e, m, p))
                     for e, m, p in ((ei, n, pos), (ei, n, None), (rnd, 300, None), (star, 11, None))]

    # ---- Unpool ----
    xc = torch.randn(5, 8, generator=gen)
    idx = torch.tensor([0, 3, 4, 8, 9])
    up = B["Unpool"]()
    out["unpool"] = dict(x=xc, indices=idx, n=12, out=up(xc, idx, 12), out3=up(xc.view(1, 5, 8), idx, 12))

    # ---- WeightedEdgeConv ----
    wec = []
    for aggr, (e_, m_, p_) in (("add", (ei, n, pos)), ("mean", (rnd, 300, torch.randn(300, 3, generator=gen)))):
        torch.manual_seed(11)
        conv = B["WeightedEdgeConv"](128, 128, aggr=aggr)
        x = torch.randn(m_, 128, generator=gen).requires_grad_(True)
        o, w = conv(x, e_, p_)
        go, gw = torch.randn(m_, 128, generator=gen), torch.randn(e_.shape[1], 1, generator=gen)
        torch.autograd.backward([o, w], [go, gw])
        rec = dict(aggr=aggr, edge_index=e_, pos=p_, x=x.detach().clone(), seed=11, state_sums=checksums(conv.state_dict()),
                   out=o.detach(), w=w.detach(), g_out=go, g_w=gw, g_x=x.grad.clone(),
                   g_params={k: p.grad.clone() for k, p in conv.named_parameters()})
        # weight reuse (the up pass): edge_weights given, compute_weights=False
        conv.zero_grad()
        x2 = x.detach().clone().requires_grad_(True)
        ew = torch.rand(e_.shape[1], 1, generator=gen).requires_grad_(True)
        o2, w2 = conv(x2, e_, p_, edge_weights=ew, compute_weights=False)
        assert w2 is ew
        o2.backward(go)
        rec.update(ew=ew.detach().clone(), out_reuse=o2.detach(), g_x_reuse=x2.grad.clone(), g_ew=ew.grad.clone(),
                   g_params_reuse={k: p.grad.clone() for k, p in conv.named_parameters() if p.grad is not None})
        wec.append(rec)
    out["wec"] = wec
    try:
        B["WeightedEdgeConv"](8, 8, aggr="max")(torch.zeros(3, 8), torch.tensor([[0], [1]]), torch.zeros(3, 2))
        out["wec_bad_aggr"] = None
    except ValueError as e:
        out["wec_bad_aggr"] = str(e)

    # ---- GMP ----
    torch.manual_seed(12)
    gmp = B["GMP"](128, 128, 128)
    x = torch.randn(300, 128, generator=gen).requires_grad_(True)          # on the random multigraph
    ea = torch.randn(rnd.shape[1], 128, generator=gen).requires_grad_(True)
    xo, eo = gmp(x, ea, rnd)
    gx, ge = torch.randn(300, 128, generator=gen), torch.randn(rnd.shape[1], 128, generator=gen)
    torch.autograd.backward([xo, eo], [gx, ge])
    out["gmp"] = dict(seed=12, state_sums=checksums(gmp.state_dict()), edge_index=rnd, x=x.detach().clone(),
                      e=ea.detach().clone(), x_out=xo.detach(), e_out=eo.detach(), g_xo=gx, g_eo=ge, g_x=x.grad.clone(),
                      g_e=ea.grad.clone(), g_params={k: p.grad.clone() for k, p in gmp.named_parameters()})
    out["gmp_silu_act"] = type(B["GMP"](8, 8, 8, activation="silu").edge_mlp[1]).__name__

    # ---- hierarchy + full model ----
    data = types.SimpleNamespace(edge_index=ei, pos=pos)
    for levels in (1, 3):
        multi = S["MultiScaleGraphPreprocessor"](num_levels=levels).create_multiscale_graph(data)
        torch.manual_seed(13)
        net = S["BSMS_MeshGraphNet"](6, 3, 4, num_levels=levels)
        na = mesh.node_attr.clone().requires_grad_(True)
        pred = net(na, mesh.edge_attr, ei, multi)
        probe = torch.randn(n, 4, generator=gen)
        (pred * probe).sum().backward()
        # parameter gradients of the full model are recorded as two scalars per parameter (norm and a seeded random
        # projection, float64) instead of 0.9 M floats
        pg = torch.Generator().manual_seed(99)
        g_params = {}
        for k, p in net.named_parameters():
            r = torch.randn(p.shape, generator=pg, dtype=torch.float64)
            if p.grad is not None:
                g_params[k] = (float(p.grad.double().norm()), float((p.grad.double() * r).sum()))
        out[f"model_L{levels}"] = dict(
            multi={k: list(v) for k, v in multi.items()}, out=pred.detach(), probe=probe, g_node=na.grad.clone(),
            g_params=g_params, seed=13, state_sums=checksums(net.state_dict()))
    try:
        net(mesh.node_attr, mesh.edge_attr, ei)
        out["model_no_multi"] = None
    except ValueError as e:
        out["model_no_multi"] = str(e)
    cfg = {"model": {"input_node_dim": 6, "input_edge_dim": 3, "output_node_dim": 4, "num_levels": 2, "hidden_dim": 128}}
    m = S["create_bsms_model_from_config"](cfg)
    out["from_config"] = dict(num_levels=m.num_levels, latent_dim=m.latent_dim, keys=sorted(m.state_dict().keys()))
    path = os.path.join(ROOT, "tests", "golden", "bistride.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    for k, v in out.items():
        print("  ", k, type(v).__name__)


if __name__ == "__main__":
    main()
