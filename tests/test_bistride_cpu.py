"""CPU checks of the BFS-bistride oracle (oracle/bistride_oracle.py) and of the host-side mirror of
`models.bistride_ops` / the older `models.bsms_mgn` design.  The reference ships these components only as
CPython 3.11 bytecode; besides the vectors recorded by executing that bytecode (tests/test_bistride_golden.py) the
oracle is checked here with known-answer cases worked by hand and an independent nn.Module evaluation of the decoded
architecture."""
import pytest
import torch
import torch.nn.functional as F
from torch import nn

from oracle import bistride_oracle as B


def _undirected(pairs):
    e = pairs + [(b, a) for a, b in pairs]
    return torch.tensor(sorted(e), dtype=torch.long).t().contiguous()


def test_bfs_known_answers():
    # path 0-1-2-3-4 plus an isolated node 5
    ei = _undirected([(0, 1), (1, 2), (2, 3), (3, 4)])
    assert B.bfs_distance(ei, 6, 0).tolist() == [0, 1, 2, 3, 4, -1]
    assert B.bfs_distance(ei, 6, 2).tolist() == [2, 1, 0, 1, 2, -1]
    # directed edges are followed sender -> receiver only
    ei = torch.tensor([[0, 1, 3], [1, 2, 2]])
    assert B.bfs_distance(ei, 4, 0).tolist() == [0, 1, 2, -1]
    assert B.bfs_distance(ei, 4, 2).tolist() == [-1, -1, 0, -1]
    # duplicates and self-loops change nothing
    ei = torch.tensor([[0, 0, 0, 1, 1], [0, 1, 1, 2, 1]])
    assert B.bfs_distance(ei, 3, 0).tolist() == [0, 1, 2]


def test_select_known_answers():
    # 3x3 grid, 4-neighbour, centre = node 4 (pos centroid): even levels = centre + corners
    pairs = [(r * 3 + c, r * 3 + c + 1) for r in range(3) for c in range(2)] + \
            [(r * 3 + c, (r + 1) * 3 + c) for r in range(2) for c in range(3)]
    ei = _undirected(pairs)
    pos = torch.tensor([[c, r] for r in range(3) for c in range(3)], dtype=torch.float32)
    assert B.select_bistride_nodes(ei, 9, pos).tolist() == [0, 2, 4, 6, 8]
    # without pos the seed is the node with the most outgoing edges (first on ties) -> also the centre here
    assert B.select_bistride_nodes(ei, 9, None).tolist() == [0, 2, 4, 6, 8]
    # star with 9 leaves: even levels = {centre} = 1 < 0.3 * 10 -> fallback keeps every reached node
    ei = _undirected([(0, k) for k in range(1, 10)])
    assert B.select_bistride_nodes(ei, 10, None).tolist() == list(range(10))
    # an unreachable node is never selected, fallback or not
    assert B.select_bistride_nodes(ei, 11, None).tolist() == list(range(10))


def test_multiscale_known_answer():
    # path 0..6, seed = node 3 (centroid): level 1 keeps {1,3,5}; no two kept nodes are adjacent -> no coarse edges;
    # level 2: seed is then the centroid of the 3 kept nodes, nobody else reachable -> 1 < 0.9 -> fallback = {seed}
    ei = _undirected([(k, k + 1) for k in range(6)])
    pos = torch.arange(7, dtype=torch.float32).view(-1, 1)
    m = B.create_multiscale_graph(ei, pos, num_levels=2)
    assert m["node_indices"][0].tolist() == [1, 3, 5]
    assert m["edge_indices"][1].shape == (2, 0)
    assert m["num_nodes"] == [7, 3, 1] and m["node_indices"][1].tolist() == [1]
    assert m["positions"][1].view(-1).tolist() == [1.0, 3.0, 5.0]
    # triangle fan keeps same-level edges: nodes 0 (centre), ring 1..4 closed: d = [0,1,1,1,1] -> only the centre is even
    ei = _undirected([(0, 1), (0, 2), (0, 3), (0, 4), (1, 2), (2, 3), (3, 4), (4, 1)])
    pos = torch.tensor([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1]], dtype=torch.float32)
    m = B.create_multiscale_graph(ei, pos, num_levels=1)
    assert m["node_indices"][0].tolist() == [0, 1, 2, 3, 4]          # 1 < 1.5 -> fallback keeps all
    assert torch.equal(m["edge_indices"][1], ei)


class _RefWEC(nn.Module):
    """The decoded architecture written as plain modules (independent of the oracle's functional form)."""

    def __init__(self, i, o):
        super().__init__()
        self.edge_weight_mlp = nn.Sequential(nn.Linear(2 * i + 1, 64), nn.ReLU(), nn.Linear(64, 1), nn.Sigmoid())
        self.transform = nn.Linear(i, o)

    def forward(self, x, ei, pos):
        s, d = ei
        w = self.edge_weight_mlp(torch.cat([x[s], x[d], (pos[d] - pos[s]).norm(dim=1, keepdim=True)], 1))
        out = torch.zeros(x.shape[0], self.transform.out_features).index_add_(0, d, self.transform(x)[s] * w)
        return out, w


def test_wec_oracle_matches_module_form():
    torch.manual_seed(0)
    m = _RefWEC(8, 12)
    x, pos = torch.randn(20, 8), torch.randn(20, 2)
    ei = torch.randint(0, 20, (2, 90))
    out, w = m(x, ei, pos)
    sd = m.state_dict()
    o2, w2 = B.wec(sd, "", x, ei, pos)
    assert torch.allclose(out, o2, atol=1e-6) and torch.allclose(w, w2, atol=1e-7)
    o3, _ = B.wec(sd, "", x, ei, pos, edge_weights=w2, compute_weights=False, aggr="mean")
    cnt = torch.bincount(ei[1], minlength=20).clamp(min=1).view(-1, 1)
    assert torch.allclose(o3, out / cnt, atol=1e-6)
    with pytest.raises(ValueError):
        B.wec(sd, "", x, ei, pos, aggr="max")


def test_gmp_oracle_is_an_mgn_step_with_two_linear_mlps():
    torch.manual_seed(1)
    D = 16
    edge_mlp = nn.Sequential(nn.Linear(3 * D, D), nn.ReLU(), nn.Linear(D, D), nn.LayerNorm(D))
    node_mlp = nn.Sequential(nn.Linear(2 * D, D), nn.ReLU(), nn.Linear(D, D), nn.LayerNorm(D))
    sd = {f"edge_mlp.{k}": v for k, v in edge_mlp.state_dict().items()}
    sd.update({f"node_mlp.{k}": v for k, v in node_mlp.state_dict().items()})
    x, e = torch.randn(30, D), torch.randn(100, D)
    ei = torch.randint(0, 30, (2, 100))
    e2 = e + edge_mlp(torch.cat([x[ei[0]], x[ei[1]], e], 1))
    x2 = x + node_mlp(torch.cat([x, torch.zeros(30, D).index_add_(0, ei[1], e2)], 1))
    xo, eo = B.gmp(sd, "", x, e, ei)
    assert torch.allclose(xo, x2, atol=1e-6) and torch.allclose(eo, e2, atol=1e-6)


def test_unpool_oracle():
    xc = torch.arange(6.0).view(3, 2)
    out = B.unpool(xc, torch.tensor([0, 2, 5]), 6)
    assert out.tolist() == [[0, 1], [0, 0], [2, 3], [0, 0], [0, 0], [4, 5]]
    assert B.unpool(xc.view(1, 3, 2), torch.tensor([0, 2, 5]), 6).shape == (1, 6, 2)


def test_host_mirror_state_dict_and_errors():
    import aero_gnn_b200.models as M
    import models.bistride_ops as shim
    import models.bsms_mgn as shim2
    assert shim.GMP is M.GMP and shim2.BSMS_MeshGraphNet is M.BSMS_MeshGraphNet
    g = M.GMP(128, 128, 128)
    assert sorted(g.state_dict()) == sorted(
        [f"{m}.{i}.{p}" for m in ("edge_mlp", "node_mlp") for i in (0, 2, 3) for p in ("weight", "bias")])
    assert g.edge_mlp[0].weight.shape == (128, 384) and g.node_mlp[0].weight.shape == (128, 256)
    w = M.WeightedEdgeConv(128, 128)
    assert sorted(w.state_dict()) == sorted(
        ["edge_weight_mlp.0.weight", "edge_weight_mlp.0.bias", "edge_weight_mlp.2.weight", "edge_weight_mlp.2.bias",
         "transform.weight", "transform.bias"])
    assert w.edge_weight_mlp[0].weight.shape == (64, 257) and w.aggr == "add"
    net = M.BSMS_MeshGraphNet(6, 3, 4, num_levels=2)
    keys = set(net.state_dict())
    assert {"bsgmp.down_gmps.2.edge_mlp.0.weight", "bsgmp.down_edge_convs.1.transform.bias",
            "bsgmp.bottom_gmp.node_mlp.3.weight", "bsgmp.up_edge_convs.0.edge_weight_mlp.2.weight",
            "node_encoder.layer_norm.weight", "decoder.layers.3.bias"} <= keys
    assert not any(k.startswith("decoder.layer_norm") for k in keys) and not any("unpools" in k for k in keys)
    assert isinstance(M.GMP(128, 128, 128, activation="silu").edge_mlp[1], nn.SiLU)
    # no CPU fallback, reference-style errors
    x = torch.zeros(4, 128)
    ei = torch.tensor([[0, 1], [1, 2]])
    with pytest.raises(RuntimeError, match="CUDA"):
        g(x, torch.zeros(2, 128), ei)
    with pytest.raises(RuntimeError, match="CUDA"):
        w(x, ei, torch.zeros(4, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        M.BistridePooling.select_bistride_nodes(ei, 4)
    cfg = {"model": {"input_node_dim": 6, "input_edge_dim": 3, "output_node_dim": 4, "num_levels": 1}}
    assert isinstance(M.create_bsms_model_from_config(cfg), M.BSMS_MeshGraphNet)
