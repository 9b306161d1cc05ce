"""GPU parity tests proper (-m gpu): every call goes through the C ABI of libaero_sm100.so and is compared
with the CPU oracle (oracle/mgn_oracle.py) and with the golden fixtures recorded from the reference.

Tolerances (north_star): integer / index work bit-exact; fp32 path <= 1e-5 relative per layer output;
bf16 path <= 1e-2 relative on final node predictions.  Gradients are checked at 1e-4 (fp32).
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err, rel_l2, rrmse
from oracle import mgn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5
GTOL = 1e-4


def _mods():
    import aero_gnn_b200.models as M
    return M


# ------------------------------------------------------------------------------------------------
# integer kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,e,seed", [(1, 0, 0), (5, 1, 1), (50, 400, 2), (37, 301, 3), (1000, 5000, 4),
                                      (3, 3000, 5), (70000, 300001, 6)])
def test_graph_plan_bit_exact(n, e, seed):
    from aero_gnn_b200 import ops
    rng = np.random.default_rng(seed)
    ei = rng.integers(0, n, size=(2, e)).astype(np.int64)
    plan = ops.build_graph_plan(torch.from_numpy(ei).to(DEV), n)
    rowptr, perm, src, dst, sptr, sperm = O.receiver_csr(ei, n)
    assert np.array_equal(plan.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(plan.perm.cpu().numpy(), perm)
    assert np.array_equal(plan.src.cpu().numpy(), src) and np.array_equal(plan.dst.cpu().numpy(), dst)
    assert np.array_equal(plan.sptr.cpu().numpy(), sptr)
    assert np.array_equal(plan.sperm.cpu().numpy(), sperm)


def test_graph_plan_rejects_out_of_range():
    from aero_gnn_b200 import ops
    ei = torch.tensor([[0, 1, 7], [1, 2, 0]], device=DEV)
    with pytest.raises(IndexError):
        ops.build_graph_plan(ei, 3)


@pytest.mark.parametrize("n,bits", [(0, 8), (1, 1), (255, 8), (4097, 13), (100003, 40), (100003, 64)])
def test_radix_sort_matches_stable_sort(n, bits):
    import ctypes as C
    from aero_gnn_b200 import lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(n + bits)
    hi = (1 << min(bits, 62)) - 1
    keys = torch.randint(0, hi + 1, (n,), generator=g, dtype=torch.int64)
    if n > 10:
        keys[n // 2:] = keys[: n - n // 2].clone()  # force ties: stability matters
    vals = torch.arange(n, dtype=torch.int32)
    ka, va = keys.to(DEV), vals.to(DEV)
    kb, vb = torch.empty_like(ka), torch.empty_like(va)
    ws = torch.empty(max(lib.aero_sort_pairs_workspace_bytes(n), 256), dtype=torch.uint8, device=DEV)
    rc = lib.aero_sort_pairs_u64(ka.data_ptr(), va.data_ptr(), kb.data_ptr(), vb.data_ptr(), n, bits, ws.data_ptr(),
                                 ws.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.aero_last_error()
    ref = torch.sort(keys, stable=True)
    assert torch.equal(kb.cpu(), ref.values) and torch.equal(vb.cpu().long(), ref.indices)


def test_bistride_indices_bit_exact_vs_golden():
    from aero_gnn_b200 import pooling
    g = load_golden("bsms")
    f2c, cb = pooling.stride_pool_assign(g["batch"].to(DEV), g["pos"][:, 0].to(DEV), 2)
    assert torch.equal(f2c.cpu(), g["l1_f2c"]) and torch.equal(cb.cpu(), g["l1_cb"])
    cei, inv, gptr, glist = pooling.coarsen_edges(g["edge_index"].to(DEV), f2c, cb.numel())
    assert torch.equal(cei.cpu(), g["l1_cei"])
    _, inv_ref = O.coarsen_edge_indices(g["edge_index"].numpy(), g["l1_f2c"].numpy(), cb.numel())
    assert np.array_equal(inv.cpu().numpy(), inv_ref)
    f2c2, cb2 = pooling.stride_pool_assign(g["l1_cb"].to(DEV), g["l1_cpos"][:, 0].to(DEV), 2)
    assert torch.equal(f2c2.cpu(), g["l2_f2c"]) and torch.equal(cb2.cpu(), g["l2_cb"])
    f2c0, cb0 = pooling.stride_pool_assign(g["batch"].to(DEV), None, 2)
    assert torch.equal(f2c0.cpu(), g["nopos_f2c"]) and torch.equal(cb0.cpu(), g["nopos_cb"])
    cei0, _, _, _ = pooling.coarsen_edges(g["edge_index"].to(DEV), f2c0, cb0.numel())
    assert torch.equal(cei0.cpu(), g["nopos_cei"])
    s3 = load_golden("bsms_stride3")
    f2c3, cb3 = pooling.stride_pool_assign(s3["batch"].to(DEV), s3["pos"][:, 0].to(DEV), 3)
    assert torch.equal(f2c3.cpu(), s3["f2c"]) and torch.equal(cb3.cpu(), s3["cb"])
    cei3, _, _, _ = pooling.coarsen_edges(s3["edge_index"].to(DEV), f2c3, cb3.numel())
    assert torch.equal(cei3.cpu(), s3["cei"])


@pytest.mark.parametrize("seed,n,graphs,stride", [(0, 1000, 1, 2), (1, 5003, 7, 2), (2, 4001, 3, 3), (3, 9, 9, 2)])
def test_bistride_indices_random_vs_oracle(seed, n, graphs, stride):
    from aero_gnn_b200 import pooling
    rng = np.random.default_rng(seed)
    batch = np.sort(rng.integers(0, graphs, size=n)).astype(np.int64) * 3 + 1   # non-trivial graph ids
    posx = rng.standard_normal(n).astype(np.float32)
    posx[rng.integers(0, n, size=n // 10)] = 0.25                              # ties -> index order
    ei = rng.integers(0, n, size=(2, 6 * n)).astype(np.int64)
    f2c_ref, cb_ref = O.stride_pool_indices(batch, posx, stride)
    f2c, cb = pooling.stride_pool_assign(torch.from_numpy(batch).to(DEV), torch.from_numpy(posx).to(DEV), stride)
    assert np.array_equal(f2c.cpu().numpy(), f2c_ref) and np.array_equal(cb.cpu().numpy(), cb_ref)
    cei_ref, inv_ref = O.coarsen_edge_indices(ei, f2c_ref, cb_ref.shape[0])
    cei, inv, gptr, glist = pooling.coarsen_edges(torch.from_numpy(ei).to(DEV), f2c, cb.numel())
    assert np.array_equal(cei.cpu().numpy(), cei_ref) and np.array_equal(inv.cpu().numpy(), inv_ref)
    # group lists: members of each coarse edge ascending
    gp, gl = gptr.cpu().numpy(), glist.cpu().numpy()
    assert gp[0] == 0 and gp[-1] == ei.shape[1]
    assert np.array_equal(inv_ref[gl], np.repeat(np.arange(cei_ref.shape[1]), np.diff(gp)))


# ------------------------------------------------------------------------------------------------
# gathers / segmented reductions
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("width", [128, 3, 2])
def test_segment_reduce_and_gather(dtype, width):
    from aero_gnn_b200 import ops
    g = torch.Generator().manual_seed(0)
    n, e = 200, 1500
    ei = torch.randint(0, n, (2, e), generator=g)
    ei[1, ei[1] == 17] = 18                                           # an empty receiver
    plan = ops.build_graph_plan(ei.to(DEV), n)
    vals = torch.randn(e, width, generator=g).to(dtype)
    v = vals.to(DEV)
    for mean in (False, True):
        out = ops.segment_reduce(v, plan.rowptr, plan.perm, n, mean=mean, out_dtype=torch.float32)
        ref = (O.scatter_mean if mean else O.scatter_add)(vals.float(), ei[1], n)
        if dtype == torch.float32 and not mean:
            assert torch.equal(out.cpu(), ref)                        # same accumulation order as CPU scatter_add_
        assert rel_err(out, ref) < 1e-6
    assert float(out[17].abs().max()) == 0.0
    got = ops.gather_rows(v, plan.perm)
    assert torch.equal(got.cpu(), vals[plan.perm.cpu().long()])
    got = ops.gather_rows(v, plan.perm, add=v)
    assert torch.equal(got.float().cpu(), (vals[plan.perm.cpu().long()].float() + vals.float()).to(dtype).float())


# ------------------------------------------------------------------------------------------------
# one processor layer
# ------------------------------------------------------------------------------------------------
def _layer_from_golden(name, dtype=torch.float32):
    M = _mods()
    g = load_golden(name)
    layer = M.MeshGraphNetLayer(128, 128, 128, **g["kwargs"])
    layer.load_state_dict(g["state"])
    return g, layer.to(DEV).to(dtype)


@pytest.mark.parametrize("name", ["layer_sum_L2_add", "layer_cat_L1_mean"])
def test_layer_forward_backward_fp32_vs_golden(name):
    g, layer = _layer_from_golden(name)
    x = g["x"].to(DEV).requires_grad_(True)
    e = g["e"].to(DEV).requires_grad_(True)
    xo, eo = layer(x, e, g["edge_index"].to(DEV))
    assert rel_err(xo, g["x_out"]) < TOL, rel_err(xo, g["x_out"])
    assert rel_err(eo, g["e_out"]) < TOL, rel_err(eo, g["e_out"])
    loss = (torch.cat([xo, eo], 0) * g["probe"].to(DEV)).sum()
    loss.backward()
    assert rel_err(x.grad, g["g_x"]) < GTOL and rel_err(e.grad, g["g_e"]) < GTOL
    for n, p in layer.named_parameters():
        assert rel_err(p.grad, g["g_params"][n]) < GTOL, (n, rel_err(p.grad, g["g_params"][n]))


@pytest.mark.parametrize("n,e", [(64, 64), (65, 129), (300, 2111), (10, 700), (129, 1)])
def test_layer_fp32_vs_oracle_ragged_sizes(n, e):
    """Tile-ragged shapes, zero-degree nodes, receivers whose edges straddle row tiles (10 nodes / 700 edges)."""
    g, layer = _layer_from_golden("layer_sum_L2_add")
    gen = torch.Generator().manual_seed(n * 1000 + e)
    x = torch.randn(n, 128, generator=gen)
    ea = torch.randn(e, 128, generator=gen)
    ei = torch.randint(0, n, (2, e), generator=gen)
    sd = g["state"]
    xr, er = O.mgn_layer(sd, "", x, ea, ei, "add")
    xo, eo = layer(x.to(DEV), ea.to(DEV), ei.to(DEV))
    assert rel_err(xo, xr) < TOL and rel_err(eo, er) < TOL


def test_layer_empty_graph():
    g, layer = _layer_from_golden("layer_sum_L2_add")
    x = torch.randn(5, 128)
    ea = torch.zeros(0, 128)
    ei = torch.zeros(2, 0, dtype=torch.long)
    xr, er = O.mgn_layer(g["state"], "", x, ea, ei, "add")
    xo, eo = layer(x.to(DEV), ea.to(DEV), ei.to(DEV))
    assert eo.shape == (0, 128) and rel_err(xo, xr) < TOL


def test_layer_is_deterministic():
    g, layer = _layer_from_golden("layer_sum_L2_add")
    gen = torch.Generator().manual_seed(5)
    n, e = 3000, 20000
    x = torch.randn(n, 128, generator=gen).to(DEV).requires_grad_(True)
    ea = torch.randn(e, 128, generator=gen).to(DEV).requires_grad_(True)
    ei = torch.randint(0, n, (2, e), generator=gen).to(DEV)
    outs = []
    for _ in range(2):
        for p in layer.parameters():
            p.grad = None
        x.grad = ea.grad = None
        xo, eo = layer(x, ea, ei)
        (xo.sum() + (eo * eo).sum()).backward()
        outs.append([xo.detach().clone(), eo.detach().clone(), x.grad.clone(), ea.grad.clone()] +
                    [p.grad.clone() for p in layer.parameters()])
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_standalone_blocks_vs_golden():
    M = _mods()
    g = load_golden("blocks")
    x, e, ei = g["x"].to(DEV), g["e"].to(DEV), g["edge_index"].to(DEV)
    eb = M.EdgeBlock(128, 128, 128, 2); eb.load_state_dict(g["state_eb"]); eb.to(DEV)
    es = M.EdgeBlockSum(128, 128, 128, 0); es.load_state_dict(g["state_es"]); es.to(DEV)
    nb = M.NodeBlock(128, 128, 128, 1, aggregation="mean"); nb.load_state_dict(g["state_nb"]); nb.to(DEV)
    assert rel_err(eb(e, x, ei), g["out_eb"]) < TOL
    assert rel_err(es(e, x, ei), g["out_es"]) < TOL
    assert rel_err(nb(x, e, ei), g["out_nb"]) < TOL
    # gradients of the standalone blocks: truth = fp64 oracle autograd; the fp32 kernels must be as close to it
    # as the reference's own fp32 arithmetic is (within 3x), or within 1e-4
    def oracle_grads(dt):
        sd_eb = {k: v.to(dt) for k, v in g["state_eb"].items()}
        sd_nb = {k: v.to(dt) for k, v in g["state_nb"].items()}
        xg, eg = g["x"].to(dt).requires_grad_(True), g["e"].to(dt).requires_grad_(True)
        ref = O.edge_block(sd_eb, "", eg, xg, g["edge_index"]).square().sum() \
            + O.node_block(sd_nb, "", xg, eg, g["edge_index"], "mean").square().sum()
        return torch.autograd.grad(ref, [xg, eg])
    gx64, ge64 = oracle_grads(torch.float64)
    gx32, ge32 = oracle_grads(torch.float32)
    xd, ed = x.clone().requires_grad_(True), e.clone().requires_grad_(True)
    (eb(ed, xd, ei).square().sum() + nb(xd, ed, ei).square().sum()).backward()
    assert rel_err(xd.grad, gx64) < max(GTOL, 3 * rel_err(gx32, gx64))
    assert rel_err(ed.grad, ge64) < max(GTOL, 3 * rel_err(ge32, ge64))
    with pytest.raises(ValueError):
        M.NodeBlock(128, 128, 128, 1, aggregation="sum").to(DEV)(x, e, ei)


# ------------------------------------------------------------------------------------------------
# models
# ------------------------------------------------------------------------------------------------
def test_mgn_model_fp32_vs_golden_with_grads():
    M = _mods()
    g = load_golden("mgn")
    net = M.MeshGraphNet(6, 3, 4, **g["kwargs"])
    net.load_state_dict(g["state"])
    net.to(DEV)
    na = g["node_attr"].to(DEV).requires_grad_(True)
    ea = g["edge_attr"].to(DEV).requires_grad_(True)
    out = net(na, ea, g["edge_index"].to(DEV))
    assert rel_err(out, g["out"]) < TOL, rel_err(out, g["out"])
    (out * g["probe"].to(DEV)).sum().backward()
    assert rel_err(na.grad, g["g_node"]) < GTOL and rel_err(ea.grad, g["g_edge"]) < GTOL
    for n, p in net.named_parameters():
        assert rel_err(p.grad, g["g_params"][n]) < GTOL, (n, rel_err(p.grad, g["g_params"][n]))


def test_bsms_model_fp32_vs_golden_with_grads():
    M = _mods()
    g = load_golden("bsms")
    net = M.BiStridedMeshGraphNet(6, 3, 4, **g["kwargs"])
    net.load_state_dict(g["state"])
    net.to(DEV)
    na = g["node_attr"].to(DEV).requires_grad_(True)
    ea = g["edge_attr"].to(DEV).requires_grad_(True)
    ei, b, p = g["edge_index"].to(DEV), g["batch"].to(DEV), g["pos"].to(DEV)
    out = net(na, ea, ei, b, p)
    assert rel_err(out, g["out"]) < TOL, rel_err(out, g["out"])
    (out * g["probe"].to(DEV)).sum().backward()
    assert rel_err(na.grad, g["g_node"]) < GTOL and rel_err(ea.grad, g["g_edge"]) < GTOL
    for n, prm in net.named_parameters():
        assert rel_err(prm.grad, g["g_params"][n]) < GTOL, (n, rel_err(prm.grad, g["g_params"][n]))
    # reference-compatible _downsample: indices exact, means within tolerance
    with torch.no_grad():
        xh, eh = net.node_encoder(na), net.edge_encoder(ea)
        cx, ce, cei, cb, cpos, f2c = net._downsample(xh, eh, ei, b, p)
    assert torch.equal(f2c.cpu(), g["l1_f2c"]) and torch.equal(cei.cpu(), g["l1_cei"]) and torch.equal(cb.cpu(), g["l1_cb"])
    assert torch.equal(cpos.cpu(), g["l1_cpos"])
    assert rel_err(cx, g["l1_cx"]) < TOL and rel_err(ce, g["l1_ce"]) < TOL
    assert torch.equal(net._unpool_nodes(cx, f2c).cpu(), cx.cpu()[g["l1_f2c"]])


def test_pool_and_fourier_models_vs_golden():
    M = _mods()
    g = load_golden("poolmgn")
    net = M.poolMGN(6, 3, 4, global_pool_method="mean", num_hidden_layers_global_encoder=2, global_dim=128, **g["kwargs"])
    net.load_state_dict(g["state"]); net.to(DEV)
    out = net(g["node_attr"].to(DEV), g["edge_attr"].to(DEV), g["edge_index"].to(DEV), g["batch"].to(DEV))
    assert rel_err(out, g["out"]) < TOL
    g = load_golden("fouriermgn")
    net = M.FourierMeshGraphNet(6, 3, 4, **g["kwargs"])
    net.load_state_dict(g["state"]); net.to(DEV)
    out = net(g["node_attr"].to(DEV), g["edge_attr"].to(DEV), g["edge_index"].to(DEV))
    assert rel_err(out, g["out"]) < TOL


def test_mgn15_airfoil_fp32_and_bf16_vs_oracle():
    """Config C1 shape (5k-node airfoil, 15 steps, config.yaml kwargs): fp32 <= 1e-5 per output; bf16 <= 1e-2
    relative error of the final node predictions (the reference's own RRMSE definition, inference.py:113-126;
    the max-norm figure is printed) against the fp32 oracle evaluated on the parameters / inputs the bf16
    model actually holds (train.py:30-33 casts the whole model to bf16; that rounding is the precision mode the
    user chose, the kernels' arithmetic error is what is bounded here)."""
    M = _mods()
    from aero_gnn_b200.meshes import airfoil_o_mesh
    kw = dict(load_golden("mgn")["kwargs"])
    kw["processor_size"] = 15
    torch.manual_seed(0)
    net = M.MeshGraphNet(6, 3, 4, **kw)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    mesh = airfoil_o_mesh(100, 50, seed=0)
    with torch.no_grad():
        ref = O.mgn_forward(sd, mesh.node_attr, mesh.edge_attr, mesh.edge_index, "add")
        net.to(DEV)
        out = net(mesh.node_attr.to(DEV), mesh.edge_attr.to(DEV), mesh.edge_index.to(DEV))
        assert rel_err(out, ref) < TOL, rel_err(out, ref)
        net16 = net.to(torch.bfloat16)
        sd16 = {k: v.detach().float().cpu() for k, v in net16.state_dict().items()}
        na16, ea16 = mesh.node_attr.to(torch.bfloat16), mesh.edge_attr.to(torch.bfloat16)
        ref16 = O.mgn_forward(sd16, na16.float(), ea16.float(), mesh.edge_index, "add")
        out16 = net16(na16.to(DEV), ea16.to(DEV), mesh.edge_index.to(DEV))
        assert out16.dtype == torch.bfloat16
        err = rrmse(out16.float(), ref16)
        print("bf16 MGN-15 RRMSE vs fp32 oracle on bf16-held parameters:", err, " max-norm:",
              rel_err(out16.float(), ref16), " RRMSE vs fp32 parameters:", rrmse(out16.float(), ref))
        assert err < 1e-2, err


def test_bf16_layer_vs_fp32_oracle_and_grads():
    g, layer = _layer_from_golden("layer_sum_L2_add", torch.bfloat16)
    ei = g["edge_index"]
    x16, e16 = g["x"].to(torch.bfloat16), g["e"].to(torch.bfloat16)
    xb = x16.to(DEV).requires_grad_(True)
    eb = e16.to(DEV).requires_grad_(True)
    xo, eo = layer(xb, eb, ei.to(DEV))
    # oracle on the bf16-rounded inputs and weights, fp32 math, autograd for the gradients
    sd = {k: v.to(torch.bfloat16).float() for k, v in g["state"].items()}
    xr_in, er_in = x16.float().requires_grad_(True), e16.float().requires_grad_(True)
    xr, er = O.mgn_layer(sd, "", xr_in, er_in, ei, "add")
    assert rel_l2(xo.float(), xr) < 1e-2 and rel_l2(eo.float(), er) < 1e-2
    gx_ref, ge_ref = torch.autograd.grad((torch.cat([xr, er], 0) * g["probe"]).sum(), [xr_in, er_in])
    (torch.cat([xo, eo], 0).float() * g["probe"].to(DEV)).sum().backward()
    # yardstick: the reference's own bf16 mode (pure bf16 autograd, train.py:30-33) against the same fp32 truth
    sd16 = {k: v.to(torch.bfloat16) for k, v in g["state"].items()}
    x16r, e16r = x16.clone().requires_grad_(True), e16.clone().requires_grad_(True)
    xq, eq = O.mgn_layer(sd16, "", x16r, e16r, ei, "add")
    gx16, ge16 = torch.autograd.grad((torch.cat([xq, eq], 0).float() * g["probe"]).sum(), [x16r, e16r])
    ex, ee = rel_l2(xb.grad.float(), gx_ref), rel_l2(eb.grad.float(), ge_ref)
    bx, be = rel_l2(gx16.float(), gx_ref), rel_l2(ge16.float(), ge_ref)
    print("bf16 layer grad rel-L2 err:", ex, ee, " reference bf16 mode:", bx, be)
    assert ex <= max(1e-2, 2.0 * bx) and ee <= max(1e-2, 2.0 * be)


# ---- standalone MLP (encoders): fused tail after the first Linear vs the plain row-wise chain (mlp.py:40-51) ----
@pytest.mark.parametrize("rows,in_dim,nh,act", [(1, 4, 1, "relu"), (127, 6, 2, "relu"), (1000, 3, 2, "tanh"),
                                                (4097, 5, 1, "elu")])
def test_mlp_fused_tail_fp32_vs_chain(rows, in_dim, nh, act):
    from aero_gnn_b200.models.mlp import MLP
    dev = torch.device("cuda", 0)
    torch.manual_seed(rows)
    mlp = MLP(in_dim, 128, 128, num_hidden_layers=nh, activation_fn=act).to(dev)
    x = torch.randn(rows, in_dim, device=dev, requires_grad=True)
    g = torch.randn(rows, 128, device=dev)
    assert mlp._fusable(x)
    out = mlp(x)
    out.backward(g)
    got = [x.grad.clone()] + [p.grad.clone() for p in mlp.parameters()]
    x.grad = None
    mlp.zero_grad(set_to_none=True)
    h = x                                           # the chain of the reference, fp64 on the same parameters
    lay = list(mlp.layers)
    h = h.double()
    for i, lin in enumerate(lay):
        h = torch.nn.functional.linear(h, lin.weight.double(), lin.bias.double())
        if i < len(lay) - 1:
            h = getattr(torch.nn.functional, act)(h)
    ref = torch.nn.functional.layer_norm(h, (128,), mlp.layer_norm.weight.double(), mlp.layer_norm.bias.double())
    ref.backward(g.double())
    want = [x.grad] + [p.grad for p in mlp.parameters()]
    assert rel_err(out, ref) < 1e-5, rel_err(out, ref)                 # fp32 tolerance of the north star
    for a, b in zip(got, want):
        assert rel_err(a, b) < 1e-4, rel_err(a, b)


def test_mlp_fused_tail_bf16_and_unfusable_shapes():
    from aero_gnn_b200.models.mlp import MLP
    dev = torch.device("cuda", 0)
    torch.manual_seed(11)
    mlp = MLP(6, 128, 128, num_hidden_layers=2).to(dev).to(torch.bfloat16)
    x = torch.randn(3000, 6, device=dev).to(torch.bfloat16)
    out = mlp(x)
    h = x.float()
    lay = list(mlp.layers)
    for i, lin in enumerate(lay):
        h = torch.nn.functional.linear(h, lin.weight.float(), lin.bias.float())
        if i < len(lay) - 1:
            h = torch.relu(h)
    ref = torch.nn.functional.layer_norm(h, (128,), mlp.layer_norm.weight.float(), mlp.layer_norm.bias.float())
    assert rrmse(out.float(), ref) < 1e-2, rrmse(out.float(), ref)     # bf16 tolerance of the north star
    # shapes the kernel does not cover stay on the library chain
    assert not MLP(128, 128, 5, num_hidden_layers=2, use_layer_norm=True).to(dev)._fusable(torch.randn(4, 128, device=dev))
    assert not MLP(6, 128, 5, num_hidden_layers=2, use_layer_norm=False).to(dev)._fusable(torch.randn(4, 6, device=dev))
    assert not MLP(6, 64, 64, num_hidden_layers=2).to(dev)._fusable(torch.randn(4, 6, device=dev))
    assert not MLP(6, 128, 128, num_hidden_layers=0).to(dev)._fusable(torch.randn(4, 6, device=dev))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_segment_reduce_into_column_blocks(dtype):
    """aero_segment_reduce_ld: two reductions fill the column blocks of one [n, 2w] matrix and equal the plain calls;
    malformed output views are rejected."""
    from aero_gnn_b200 import ops
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(7)
    n, rows, w = 301, 4000, 128
    src = torch.randn(rows, w, generator=g).to(dev, dtype)
    seg = torch.sort(torch.randint(0, n, (rows,), generator=g)).values
    ptr = torch.zeros(n + 1, dtype=torch.int32)
    ptr[1:] = torch.cumsum(torch.bincount(seg, minlength=n), 0).to(torch.int32)
    ptr = ptr.to(dev)
    lst = torch.randperm(rows, generator=g).to(torch.int32).to(dev)
    both = torch.full((n, 2 * w), 7.0, dtype=dtype, device=dev)
    ops.segment_reduce(src, ptr, lst, n, out=both[:, :w])
    ops.segment_reduce(src, ptr, None, n, out=both[:, w:])
    assert torch.equal(both[:, :w], ops.segment_reduce(src, ptr, lst, n))
    assert torch.equal(both[:, w:], ops.segment_reduce(src, ptr, None, n))
    with pytest.raises(RuntimeError):
        ops.segment_reduce(src, ptr, None, n, out=both[:, ::2])                  # non-unit column stride
    with pytest.raises(RuntimeError):
        ops.segment_reduce(src, ptr, None, n, out=torch.empty(n + 1, w, dtype=dtype, device=dev))


@pytest.mark.parametrize("rows,out_dim,nh,ln", [(1, 5, 2, False), (777, 4, 1, False), (3000, 128, 2, True)])
def test_mlp_fused_decoder_fp32_vs_chain(rows, out_dim, nh, ln):
    """A 128-wide input (the decoder, mgn.py:130) runs whole on the block kernel; a narrow last Linear is zero-padded."""
    from aero_gnn_b200.models.mlp import MLP
    dev = torch.device("cuda", 0)
    torch.manual_seed(rows + out_dim)
    mlp = MLP(128, 128, out_dim, num_hidden_layers=nh, use_layer_norm=ln).to(dev)
    x = torch.randn(rows, 128, device=dev, requires_grad=True)
    g = torch.randn(rows, out_dim, device=dev)
    assert mlp._fusable(x)
    out = mlp(x)
    assert out.shape == (rows, out_dim)
    out.backward(g)
    got = [x.grad.clone()] + [p.grad.clone() for p in mlp.parameters()]
    x.grad = None
    mlp.zero_grad(set_to_none=True)
    h = x.double()
    lay = list(mlp.layers)
    for i, lin in enumerate(lay):
        h = torch.nn.functional.linear(h, lin.weight.double(), lin.bias.double())
        if i < len(lay) - 1:
            h = torch.relu(h)
    if ln:
        h = torch.nn.functional.layer_norm(h, (out_dim,), mlp.layer_norm.weight.double(), mlp.layer_norm.bias.double())
    h.backward(g.double())
    want = [x.grad] + [p.grad for p in mlp.parameters()]
    assert rel_err(out, h) < 1e-5, rel_err(out, h)
    for a, b in zip(got, want):
        assert rel_err(a, b) < 1e-4, rel_err(a, b)


# ---- one-launch parameter packing (aero_multi_copy) vs the torch-op packing and its autograd ----------------------
@pytest.mark.parametrize("concat,L,dtype", [(True, 2, torch.bfloat16), (False, 1, torch.float32), (True, 0, torch.float32)])
def test_pack_step_matches_torch_packing_and_routes_gradients(concat, L, dtype):
    from aero_gnn_b200 import processor as P
    M = _mods()
    torch.manual_seed(L)
    if L == 0:
        layer = M.GMP(128, 128, 128).to(DEV).to(dtype)
        nd = 128
        e0, e2, eln = layer.edge_mlp[0], layer.edge_mlp[2], layer.edge_mlp[3]
        n0, n2, nln = layer.node_mlp[0], layer.node_mlp[2], layer.node_mlp[3]
        parts = lambda: ((e0.weight[:, 2 * nd:], [], e2.weight, e2.bias, eln.weight, eln.bias),
                         (n0.weight[:, nd:], [], n2.weight, n2.bias, nln.weight, nln.bias),
                         [e0.weight[:, :nd], e0.weight[:, nd:2 * nd], n0.weight[:, :nd]], [None, e0.bias, n0.bias])
    else:
        layer = M.MeshGraphNetLayer(128, 128, 128, L, L, do_concat_trick=concat).to(DEV).to(dtype)

        def parts():
            ep, np_ = layer.edge_block.fused_parts(), layer.node_block.fused_parts()
            return ((ep["w_e"], ep["hidden"], ep["w_out"], ep["b_out"], ep["gamma"], ep["beta"]),
                    (np_["w_a"], np_["hidden"], np_["w_out"], np_["b_out"], np_["gamma"], np_["beta"]),
                    [ep["w_s"], ep["w_d"], np_["w_x"]], [None, ep["b0"], np_["b0"]])
    sw = layer.step_weights(dtype)
    edge, node, pw, pb = parts()
    ref = [P.pack_block(*edge), P.pack_block(*node), torch.cat(pw, dim=0).to(dtype),
           torch.cat([torch.zeros_like(pb[1]), pb[1], pb[2]]).to(dtype)]
    got = [sw.w_edge, sw.w_node, sw.w_proj, sw.b_proj]
    for a, b in zip(got, ref):
        assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)
    g = torch.Generator().manual_seed(3)
    gs = [torch.randn(t.shape, generator=g).to(DEV, t.dtype) for t in got]
    torch.autograd.backward(got, gs)
    mine = {k: p.grad.clone() for k, p in layer.named_parameters()}
    layer.zero_grad(set_to_none=True)
    torch.autograd.backward(ref, gs)
    for k, p in layer.named_parameters():
        assert mine[k].dtype == p.grad.dtype and torch.equal(mine[k], p.grad), k


def test_mlp_module_vs_golden_on_gpu():
    """models/mlp.py:40-51 with widths the fused kernel does not cover (7 -> 32 -> 32 -> 16): the library chain on the
    GPU against the output recorded from the reference."""
    M = _mods()
    g = load_golden("mlp")
    m = M.MLP(7, 32, 16, 2, "relu")
    m.load_state_dict(g["state"], strict=True)
    m = m.to(DEV)
    assert not m._fusable(g["x"].to(DEV))
    assert rel_err(m(g["x"].to(DEV)), g["out"]) < TOL
