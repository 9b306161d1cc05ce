"""CPU-side checks: the C-ABI library loads and exports every declared symbol, module state_dict keys match
the reference's, error behaviour, mesh generators.  No compute call is made (no GPU here)."""
import os
import re

import pytest
import torch

from conftest import ROOT, load_golden


def test_library_exports_every_declared_symbol():
    from aero_gnn_b200 import lib
    hdr = open(os.path.join(ROOT, "include", "aero_gnn.h")).read()
    declared = set(re.findall(r"\b(aero_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    handle = lib.load()
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in include/aero_gnn.h but not exported"
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    assert handle.aero_version() >= 100
    assert handle.aero_graph_plan_workspace_bytes(1000, 100) > 0


def test_block_desc_matches_header_layout():
    import ctypes as C
    from aero_gnn_b200 import lib
    d = lib.BlockDesc
    assert d.rows.offset == 32 and d.main.offset == 72 and C.sizeof(d) == 72 + 18 * 8 and d.h0.offset == 72 + 17 * 8


def test_state_dict_keys_match_reference():
    import aero_gnn_b200.models as M
    g = load_golden("mgn")
    net = M.MeshGraphNet(6, 3, 4, **g["kwargs"])
    net.load_state_dict(g["state"], strict=True)
    g = load_golden("bsms")
    net = M.BiStridedMeshGraphNet(6, 3, 4, **g["kwargs"])
    net.load_state_dict(g["state"], strict=True)
    assert [len(b) for b in net.down_layers] == [1, 1] and len(net.bottleneck_layers) == 1
    g = load_golden("poolmgn")
    M.poolMGN(6, 3, 4, global_pool_method="mean", num_hidden_layers_global_encoder=2, global_dim=128,
              **g["kwargs"]).load_state_dict(g["state"], strict=True)
    g = load_golden("fouriermgn")
    M.FourierMeshGraphNet(6, 3, 4, **g["kwargs"]).load_state_dict(g["state"], strict=True)
    g = load_golden("blocks")
    M.EdgeBlock(128, 128, 128, 2).load_state_dict(g["state_eb"], strict=True)
    M.EdgeBlockSum(128, 128, 128, 0).load_state_dict(g["state_es"], strict=True)
    M.NodeBlock(128, 128, 128, 1).load_state_dict(g["state_nb"], strict=True)
    g = load_golden("layer_cat_L1_mean")
    M.MeshGraphNetLayer(128, 128, 128, **g["kwargs"]).load_state_dict(g["state"], strict=True)


def test_default_bsms_layer_budget():
    import aero_gnn_b200.models as M
    net = M.BiStridedMeshGraphNet(6, 3, 4, num_scales=4, layers_per_scale=2)
    assert [len(b) for b in net.down_layers] == [2, 2, 2] and len(net.bottleneck_layers) == 3
    net = M.BiStridedMeshGraphNet(6, 3, 4)   # config default: 3 scales -> [2,2], bottleneck 7
    assert len(net.bottleneck_layers) == 7


def test_error_conventions():
    import aero_gnn_b200.models as M
    with pytest.raises(ValueError):
        M.BiStridedMeshGraphNet(6, 3, 4, num_scales=0)
    with pytest.raises(ValueError):
        M.BiStridedMeshGraphNet(6, 3, 4, stride=0)
    with pytest.raises(ValueError):
        M.BiStridedMeshGraphNet(6, 3, 4, num_scales=3, layers_per_scale=[1])
    with pytest.raises(AttributeError):
        M.MLP(4, 8, 4, activation_fn="not_an_activation")
    with pytest.raises(ValueError):
        M.poolMGN(6, 3, 4, global_pool_method="median")
    nb = M.NodeBlock(128, 128, 128, aggregation="sum")      # reference default 'sum' is rejected at forward
    with pytest.raises(ValueError):
        nb.check_aggregation()
    # no CPU fallback: CPU tensors are refused loudly
    net = M.MeshGraphNet(6, 3, 4, processor_size=1, aggregation="add")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(3, 6), torch.zeros(2, 3), torch.zeros(2, 2, dtype=torch.long))


def test_mlp_module_matches_oracle():
    import aero_gnn_b200.models as M
    from oracle import mgn_oracle as O
    g = load_golden("mlp")
    m = M.MLP(7, 32, 16, 2, "relu")
    m.load_state_dict(g["state"])
    assert torch.allclose(m(g["x"]), g["out"], rtol=1e-6, atol=1e-6)
    assert torch.allclose(O.mlp(g["state"], "", g["x"]), g["out"], rtol=1e-6, atol=1e-6)
    single = M.MLP(5, 9, 3, 0)
    assert len(single.layers) == 1 and single.layers[0].weight.shape == (3, 5)


def test_packed_weight_layout():
    import aero_gnn_b200.models as M
    from aero_gnn_b200.ops import packed_floats
    layer = M.MeshGraphNetLayer(128, 128, 128, 2, 2, do_concat_trick=True)
    sw = layer.step_weights(torch.float32)
    assert sw.w_edge.numel() == packed_floats(2) and sw.w_node.numel() == packed_floats(2)
    assert sw.w_proj.shape == (384, 128) and sw.b_proj.shape == (384,)
    assert torch.equal(sw.w_edge[:16384].view(128, 128), layer.edge_block.edge_lin.detach())
    assert torch.equal(sw.w_proj[256:], layer.node_block.mlp.layers[0].weight[:, :128].detach())
    assert torch.equal(sw.w_node[:16384].view(128, 128), layer.node_block.mlp.layers[0].weight[:, 128:].detach())
    cat = M.MeshGraphNetLayer(128, 128, 128, 1, 1, do_concat_trick=False)
    sw2 = cat.step_weights(torch.float32)
    w0 = cat.edge_block.mlp.layers[0].weight.detach()
    assert torch.equal(sw2.w_edge[:16384].view(128, 128), w0[:, :128]) and torch.equal(sw2.w_proj[:128], w0[:, 128:256])


def test_synthetic_meshes():
    from aero_gnn_b200.meshes import airfoil_o_mesh, batch_meshes
    m = airfoil_o_mesh(100, 50)
    assert m.num_nodes == 5000 and m.num_edges == 29600
    key = m.edge_index[0] * m.num_nodes + m.edge_index[1]
    assert bool((key[1:] > key[:-1]).all())                          # coalesced, sorted by (sender, receiver)
    rev = m.edge_index[1] * m.num_nodes + m.edge_index[0]
    assert torch.equal(torch.sort(rev).values, key)                  # symmetric
    assert torch.unique(m.pos[:, 0]).numel() == m.num_nodes          # no x ties
    b = batch_meshes([airfoil_o_mesh(10, 5, seed=s) for s in range(3)])
    assert b.num_nodes == 150 and int(b.batch.max()) == 2
