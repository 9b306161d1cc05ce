"""oracle/bistride_oracle.py (and the host mirror's parameter initialisation) pinned against tests/golden/bistride.pt,
which oracle/gen_bistride_golden.py recorded by EXECUTING THE REFERENCE'S OWN 3.11 BYTECODE through
oracle/pyc311_vm.py: BFS distances, node selection, hierarchy and Unpool exact; WeightedEdgeConv, GMP and
BSMS_MeshGraphNet outputs <= 1e-5 relative, autograd gradients <= 1e-4."""
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import bistride_oracle as B


@pytest.fixture(scope="module")
def G():
    return load_golden("bistride")


def _state(module, sums):
    """state_dict of a mirror module built right after torch.manual_seed(seed); the reference drew the same numbers."""
    sd = {k: v.detach().clone() for k, v in module.state_dict().items()}
    assert list(sd) == list(sums)
    for k, (s, nrm) in sums.items():
        assert abs(float(sd[k].double().sum()) - s) <= 1e-9 * max(1.0, abs(s)), k
        assert abs(float(sd[k].double().norm()) - nrm) <= 1e-9 * max(1.0, nrm), k
    return sd


def test_bfs_select_unpool_exact(G):
    for r in G["bfs"]:
        assert torch.equal(B.bfs_distance(r["edge_index"], r["n"], r["start"]), r["dist"])
    assert any(int((r["dist"] < 0).sum()) > 0 for r in G["bfs"])                 # unreachable nodes are covered
    for r in G["select"]:
        assert torch.equal(B.select_bistride_nodes(r["edge_index"], r["n"], r["pos"]), r["selected"])
    assert G["select"][3]["selected"].tolist() == list(range(10))                # star: the 30 % fallback fired
    u = G["unpool"]
    assert torch.equal(B.unpool(u["x"], u["indices"], u["n"]), u["out"])
    assert torch.equal(B.unpool(u["x"].view(1, 5, 8), u["indices"], u["n"]), u["out3"])


def test_hierarchy_exact(G):
    m = G["mesh"]
    for levels in (1, 3):
        ref = G[f"model_L{levels}"]["multi"]
        got = B.create_multiscale_graph(m["edge_index"], m["pos"], levels)
        assert got["num_nodes"] == [int(v) for v in ref["num_nodes"]]
        for key in ("node_indices", "edge_indices", "positions"):
            assert len(got[key]) == len(ref[key])
            for a, b in zip(got[key], ref[key]):
                assert a.dtype == b.dtype and torch.equal(a, b), key


def test_weighted_edge_conv_vs_reference(G):
    import aero_gnn_b200.models as M
    for r in G["wec"]:
        torch.manual_seed(r["seed"])
        sd = _state(M.WeightedEdgeConv(128, 128, aggr=r["aggr"]), r["state_sums"])
        sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        x = r["x"].clone().requires_grad_(True)
        out, w = B.wec(sdr, "", x, r["edge_index"], r["pos"], aggr=r["aggr"])
        assert rel_err(out, r["out"]) < 1e-5 and rel_err(w, r["w"]) < 1e-5
        torch.autograd.backward([out, w], [r["g_out"], r["g_w"]])
        assert rel_err(x.grad, r["g_x"]) < 1e-4
        for k, g in r["g_params"].items():
            assert rel_err(sdr[k].grad, g) < 1e-4, k
        sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        x, ew = r["x"].clone().requires_grad_(True), r["ew"].clone().requires_grad_(True)
        out, _ = B.wec(sdr, "", x, r["edge_index"], r["pos"], edge_weights=ew, compute_weights=False, aggr=r["aggr"])
        assert rel_err(out, r["out_reuse"]) < 1e-5
        out.backward(r["g_out"])
        assert rel_err(x.grad, r["g_x_reuse"]) < 1e-4 and rel_err(ew.grad, r["g_ew"]) < 1e-4
        assert sorted(r["g_params_reuse"]) == ["transform.bias", "transform.weight"]
        for k, g in r["g_params_reuse"].items():
            assert rel_err(sdr[k].grad, g) < 1e-4, k
    with pytest.raises(ValueError) as e:
        B.wec(sd, "", r["x"], r["edge_index"], r["pos"], aggr="max")
    assert str(e.value) == G["wec_bad_aggr"] == "Unknown aggregation: max"


def test_gmp_vs_reference(G):
    import aero_gnn_b200.models as M
    r = G["gmp"]
    torch.manual_seed(r["seed"])
    sd = _state(M.GMP(128, 128, 128), r["state_sums"])
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x, e = r["x"].clone().requires_grad_(True), r["e"].clone().requires_grad_(True)
    xo, eo = B.gmp(sdr, "", x, e, r["edge_index"])
    assert rel_err(xo, r["x_out"]) < 1e-5 and rel_err(eo, r["e_out"]) < 1e-5
    torch.autograd.backward([xo, eo], [r["g_xo"], r["g_eo"]])
    assert rel_err(x.grad, r["g_x"]) < 1e-4 and rel_err(e.grad, r["g_e"]) < 1e-4
    for k, g in r["g_params"].items():
        assert rel_err(sdr[k].grad, g) < 1e-4, k
    assert G["gmp_silu_act"] == "SiLU" and type(M.GMP(8, 8, 8, activation="silu").edge_mlp[1]).__name__ == "SiLU"


@pytest.mark.parametrize("levels", [1, 3])
def test_bsms_meshgraphnet_vs_reference(G, levels):
    import aero_gnn_b200.models as M
    m, r = G["mesh"], G[f"model_L{levels}"]
    torch.manual_seed(r["seed"])
    sd = _state(M.BSMS_MeshGraphNet(6, 3, 4, num_levels=levels), r["state_sums"])
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    na = m["node_attr"].clone().requires_grad_(True)
    multi = B.create_multiscale_graph(m["edge_index"], m["pos"], levels)
    out = B.bsms_meshgraphnet(sdr, levels, na, m["edge_attr"], multi)
    assert rel_err(out, r["out"]) < 1e-5
    (out * r["probe"]).sum().backward()
    assert rel_err(na.grad, r["g_node"]) < 1e-4
    pg = torch.Generator().manual_seed(99)
    seen = 0
    for k, v in sd.items():
        rnd = torch.randn(v.shape, generator=pg, dtype=torch.float64)      # same stream as the generator script
        if k not in r["g_params"]:
            assert sdr[k].grad is None, k          # built but never called by the reference either
            continue
        nrm, proj = r["g_params"][k]
        g = sdr[k].grad.double()
        assert abs(float(g.norm()) - nrm) <= 1e-4 * max(nrm, 1e-6), k
        assert abs(float((g * rnd).sum()) - proj) <= 1e-4 * max(nrm * float(rnd.norm()), 1e-6), k
        seen += 1
    assert seen > 40


def test_messages_and_config_helper(G):
    import aero_gnn_b200.models as M
    assert G["model_no_multi"] == "multi_data must be provided. Use MultiScaleGraphPreprocessor to preprocess graphs before training."
    cfg = {"model": {"input_node_dim": 6, "input_edge_dim": 3, "output_node_dim": 4, "num_levels": 2, "hidden_dim": 128}}
    net = M.create_bsms_model_from_config(cfg)
    assert (net.num_levels, net.latent_dim) == (G["from_config"]["num_levels"], G["from_config"]["latent_dim"])
    assert sorted(net.state_dict().keys()) == G["from_config"]["keys"]


@pytest.mark.skipif(not __import__("os").path.exists("/root/reference/models/__pycache__/bistride_ops.cpython-311.pyc"),
                    reason="needs the reference checkout (only the container that generated the fixtures has it)")
def test_fixture_reproduces_from_the_reference_bytecode(G):
    """Where the reference checkout exists, re-run part of the generator: the bytecode interpreter must reproduce the
    committed vectors (guards the interpreter and the fixture against drifting apart)."""
    import sys
    saved = {k: sys.modules.get(k) for k in ("models", "models.mlp", "torch_scatter", "torch_geometric", "torch_geometric.nn")}
    saved_path = list(sys.path)
    try:
        from oracle import gen_bistride_golden as gen
        Bm, Sm = gen.load_reference_modules()
        for r in G["bfs"][:2] + G["bfs"][3:]:
            assert torch.equal(Bm["BistridePooling"].bfs_distance(r["edge_index"], r["n"], r["start"]), r["dist"])
        r = G["wec"][0]
        torch.manual_seed(r["seed"])
        conv = Bm["WeightedEdgeConv"](128, 128, aggr=r["aggr"])
        out, w = conv(r["x"], r["edge_index"], r["pos"])
        assert torch.equal(out, r["out"]) and torch.equal(w, r["w"])
        import types
        m = G["mesh"]
        multi = Sm["MultiScaleGraphPreprocessor"](num_levels=3).create_multiscale_graph(
            types.SimpleNamespace(edge_index=m["edge_index"], pos=m["pos"]))
        for a, b in zip(multi["node_indices"], G["model_L3"]["multi"]["node_indices"]):
            assert torch.equal(a, b)
    finally:
        sys.path[:] = saved_path
        for k, v in saved.items():          # the stand-ins must not leak into the other tests of this process
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
