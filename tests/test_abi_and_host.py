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
    assert d.rows.offset == 32 and d.main.offset == 72 and C.sizeof(d) == 72 + 21 * 8 and d.h0.offset == 72 + 17 * 8
    assert d.main_lat.offset == 72 + 18 * 8 and d.flags.offset == 28


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
    m.load_state_dict(g["state"], strict=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(g["x"])                                        # the module itself runs on the GPU only (tests/test_gpu_parity.py)
    assert torch.allclose(O.mlp(g["state"], "", g["x"]), g["out"], rtol=1e-6, atol=1e-6)
    single = M.MLP(5, 9, 3, 0)
    assert len(single.layers) == 1 and single.layers[0].weight.shape == (3, 5)


def _emulate_pack(spec, bases):
    """What aero_multi_copy does with a PackSpec, on CPU tensors (test-side emulation of the segment list)."""
    outs = [torch.full((n,), float("nan"), dtype=torch.float32) for n, _ in spec.outs]
    for bi, boff, r, c, ld, oi, ooff, old in spec.segs:
        for i in range(r):
            row = torch.zeros(c) if bi is None else bases[bi].detach().reshape(-1)[boff + i * ld: boff + i * ld + c].float()
            outs[oi][ooff + i * old: ooff + i * old + c] = row
    return outs


def test_packed_weight_layout(monkeypatch):
    """The segment list of the one-launch parameter packing (processor.pack_step) reproduces the packed layout of
    include/aero_gnn.h for both edge-block forms; every parameter element is used exactly once."""
    import aero_gnn_b200.models as M
    from aero_gnn_b200 import processor as P
    from aero_gnn_b200.ops import packed_floats
    seen = {}

    class _Stop(Exception):
        pass

    def fake_apply(spec, *bases):
        seen["spec"], seen["bases"] = spec, bases
        raise _Stop

    monkeypatch.setattr(P.PackStepFn, "apply", staticmethod(fake_apply))

    def spec_of(layer):
        with pytest.raises(_Stop):
            layer.step_weights(torch.float32)
        spec, bases = seen["spec"], seen["bases"]
        used = [0] * len(bases)
        for bi, _, r, c, *_rest in spec.segs:
            if bi is not None:
                used[bi] += r * c
        assert used == [b.numel() for b in bases]
        assert {id(b) for b in bases} == {id(p) for p in layer.parameters()}
        return _emulate_pack(spec, bases)

    layer = M.MeshGraphNetLayer(128, 128, 128, 2, 2, do_concat_trick=True)
    w_edge, w_node, w_proj, b_proj = spec_of(layer)
    assert w_edge.numel() == packed_floats(2) and w_node.numel() == packed_floats(2)
    assert not torch.isnan(torch.cat([w_edge, w_node, w_proj, b_proj])).any()
    w_proj = w_proj.view(384, 128)
    assert b_proj.shape == (384,) and torch.equal(b_proj[:128], torch.zeros(128))
    assert torch.equal(w_edge[:16384].view(128, 128), layer.edge_block.edge_lin.detach())
    assert torch.equal(w_proj[256:], layer.node_block.mlp.layers[0].weight[:, :128].detach())
    assert torch.equal(w_node[:16384].view(128, 128), layer.node_block.mlp.layers[0].weight[:, 128:].detach())
    # against the torch-op packing the single-block paths still use
    ep = layer.edge_block.fused_parts()
    ref = P.pack_block(ep["w_e"], ep["hidden"], ep["w_out"], ep["b_out"], ep["gamma"], ep["beta"]).detach()
    assert torch.equal(w_edge, ref)
    assert torch.equal(b_proj[128:256], layer.edge_block.bias.detach())
    cat = M.MeshGraphNetLayer(128, 128, 128, 1, 1, do_concat_trick=False)
    w_edge2, _, w_proj2, _ = spec_of(cat)
    w0 = cat.edge_block.mlp.layers[0].weight.detach()
    assert torch.equal(w_edge2[:16384].view(128, 128), w0[:, :128])
    assert torch.equal(w_proj2.view(384, 128)[:128], w0[:, 128:256])
    gmp = M.GMP(128, 128, 128)
    w_edge3, w_node3, w_proj3, b_proj3 = spec_of(gmp)
    e0 = gmp.edge_mlp[0].weight.detach()      # [x_src | x_dst | e]
    assert torch.equal(w_edge3[:16384].view(128, 128), e0[:, 256:]) and torch.equal(w_proj3.view(384, 128)[:128], e0[:, :128])
    assert w_edge3.numel() == packed_floats(0)


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


def test_pack_spec_rejects_what_the_copy_kernel_cannot_address():
    """processor._locate / PackSpec: column slices of contiguous matrices are fine, transposed or strided pieces and
    wrongly sized pieces are refused before anything is launched."""
    from aero_gnn_b200 import processor as P
    w = torch.nn.Parameter(torch.randn(128, 384))
    base, off, r, c, ld = P._locate(w[:, 128:256])
    assert base is w and (off, r, c, ld) == (128, 128, 128, 384)
    b = torch.nn.Parameter(torch.randn(128))
    assert P._locate(b)[1:] == (0, 1, 128, 128)
    with pytest.raises(RuntimeError, match="contiguous"):
        P._locate(w.t()[:128])
    with pytest.raises(RuntimeError, match="contiguous"):
        P._locate(w[:, ::2])
    spec = P.PackSpec()
    o = spec.out(128 * 128, torch.float32)
    spec.put(w[:, :128], o, 0, 128, 128)
    spec.put(w[:, 128:256], o, 0, 128, 128)
    assert len(spec.bases) == 1 and [s[1] for s in spec.segs] == [0, 128]      # one autograd input, two pieces
    with pytest.raises(RuntimeError, match="expected"):
        spec.put(b, o, 0, 128, 128)
    spec.put(None, o, 0)                                                          # zero fill: no base
    assert spec.segs[-1][0] is None


def test_pack_spec_is_cached_per_module_and_invalidated(monkeypatch):
    import aero_gnn_b200.models as M
    from aero_gnn_b200 import processor as P
    seen = []
    monkeypatch.setattr(P.PackStepFn, "apply", staticmethod(lambda spec, *bases: (seen.append(spec), [torch.zeros(128)] * 4)[1]))
    layer = M.MeshGraphNetLayer(128, 128, 128, 2, 2, do_concat_trick=True)
    layer.step_weights(torch.float32)
    layer.step_weights(torch.float32)
    assert seen[0] is seen[1]                                  # same segment list, no views rebuilt
    with torch.no_grad():
        layer.edge_block.bias.add_(1.0)                        # values are read at launch: still valid
    layer.step_weights(torch.float32)
    assert seen[2] is seen[0]
    layer.step_weights(torch.bfloat16)                         # other latent dtype: other outputs
    assert seen[3] is not seen[0]
    layer.edge_block.bias = torch.nn.Parameter(torch.zeros(128))   # a replaced Parameter object
    layer.step_weights(torch.bfloat16)
    assert seen[4] is not seen[3] and any(b is layer.edge_block.bias for b in seen[4].bases)
    layer.to(torch.float64)                                    # same objects, new dtype
    layer.step_weights(torch.bfloat16)
    assert seen[5] is not seen[4]
    import copy
    twin = copy.deepcopy(layer)
    twin.step_weights(torch.bfloat16)
    assert all(any(b is p for p in twin.parameters()) for b in seen[6].bases)


def test_probe_library_is_separate_and_exports_its_header():
    """The hardware probes live in libaero_probe.so (include/aero_gnn_debug.h), not in the product library."""
    from aero_gnn_b200 import lib
    hdr = open(os.path.join(ROOT, "include", "aero_gnn_debug.h")).read()
    declared = set(re.findall(r"\b(aero_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(lib.PROBE_SIGNATURES)
    product = lib.load()
    probe = lib.load_probe()
    for name in declared:
        assert hasattr(probe, name) and not hasattr(product, name), name


def test_new_entry_points_validate_arguments_before_touching_the_gpu():
    """aero_wgrad / aero_row_gemm / aero_thin_linear_*: bad arguments are refused with an error code and a message,
    and an empty problem is a no-op -- all decided on the host, so this runs without a GPU."""
    import ctypes as C
    from aero_gnn_b200 import lib as L
    lib = L.load()
    one = C.c_void_p(16)                                        # any non-null, 16-byte aligned address: never dereferenced
    assert lib.aero_wgrad_workspace_bytes(1000, 2, 1) > lib.aero_wgrad_workspace_bytes(1000, 1, 0) > 0
    rc = lib.aero_wgrad(one, 128, 3, one, 10, one, None, None, 0, None, 0, one, 1 << 30, None)     # a_panels = 3
    assert rc != 0 and b"aero_wgrad" in lib.aero_last_error()
    rc = lib.aero_wgrad(one, 64, 1, one, 10, one, None, None, 0, None, 0, one, 1 << 30, None)      # lda < 128
    assert rc != 0
    rc = lib.aero_wgrad(one, 128, 1, one, 10, one, None, None, 0, None, 0, one, 8, None)           # workspace too small
    assert rc != 0 and b"workspace" in lib.aero_last_error()
    ptrs, lds = (C.c_void_p * 3)(16, 16, 16), (C.c_int64 * 3)(128, 128, 128)
    rc = lib.aero_row_gemm(ptrs, lds, 2, one, 1, 2, None, None, 0, one, 256, 10, None)             # na and nb both > 1
    assert rc != 0 and b"aero_row_gemm" in lib.aero_last_error()
    rc = lib.aero_row_gemm(ptrs, lds, 1, one, 0, 3, None, None, 0, one, 128, 10, None)             # out_ld < 128 nb
    assert rc != 0
    assert lib.aero_row_gemm(ptrs, lds, 1, one, 0, 3, None, None, 0, one, 384, 0, None) == 0       # no rows: nothing to do
    rc = lib.aero_thin_linear_fwd(one, 4, one, None, one, 10, 17, L.AERO_BF16, None)               # K > 16
    assert rc != 0 and b"aero_thin_linear_fwd" in lib.aero_last_error()
    rc = lib.aero_thin_linear_fwd(one, 2, one, None, one, 10, 4, L.AERO_BF16, None)                # ldx < K
    assert rc != 0
    assert lib.aero_thin_linear_fwd(one, 4, one, None, one, 0, 4, L.AERO_BF16, None) == 0
    assert lib.aero_thin_linear_workspace_bytes(100000, 6) >= 128 * 7 * 4
