"""tcgen05 (UMMA) path: primitive GEMM self-test in the three operand orientations, and the fused bf16 block
kernels against the CUDA-core fp32-math kernels and the CPU oracle."""
import ctypes as C
import os

import pytest
import torch

from conftest import load_golden, rel_err, rel_l2
from oracle import mgn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_umma_gemm_selftest(mode):
    from aero_gnn_b200 import lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(mode)
    a = torch.randn(128, 128, generator=g).to(torch.bfloat16)
    b = torch.randn(128, 128, generator=g).to(torch.bfloat16)
    ad, bd = a.to(DEV), b.to(DEV)
    c = torch.zeros(128, 128, device=DEV)
    rc = lib.aero_umma_selftest(ad.data_ptr(), bd.data_ptr(), c.data_ptr(), mode,
                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.aero_last_error()
    torch.cuda.synchronize()
    af, bf = a.double(), b.double()
    ref = [af @ bf.t(), af @ bf, af.t() @ bf][mode]
    err = rel_err(c, ref)
    assert err < 1e-5, (mode, err)       # exact bf16 products, fp32 accumulation


def _run_layer(layer, x, e, ei, force_simt):
    old = os.environ.get("AERO_FORCE_SIMT")
    os.environ["AERO_FORCE_SIMT"] = "1" if force_simt else "0"
    try:
        with torch.no_grad():
            return layer(x, e, ei)
    finally:
        if old is None:
            os.environ.pop("AERO_FORCE_SIMT", None)
        else:
            os.environ["AERO_FORCE_SIMT"] = old


@pytest.mark.parametrize("name", ["layer_sum_L2_add", "layer_cat_L1_mean"])
@pytest.mark.parametrize("n,e", [(37, 301), (300, 2111), (10, 700), (129, 1), (5000, 29600)])
def test_umma_forward_matches_simt_and_oracle(name, n, e):
    import aero_gnn_b200.models as M
    from aero_gnn_b200 import lib as L
    assert L.load().aero_has_umma() == 1
    g = load_golden(name)
    layer = M.MeshGraphNetLayer(128, 128, 128, **g["kwargs"])
    layer.load_state_dict(g["state"])
    layer = layer.to(DEV).to(torch.bfloat16)
    gen = torch.Generator().manual_seed(n + e)
    x = torch.randn(n, 128, generator=gen).to(torch.bfloat16)
    ea = torch.randn(e, 128, generator=gen).to(torch.bfloat16)
    ei = torch.randint(0, n, (2, e), generator=gen)
    xs, es = _run_layer(layer, x.to(DEV), ea.to(DEV), ei.to(DEV), force_simt=True)
    xu, eu = _run_layer(layer, x.to(DEV), ea.to(DEV), ei.to(DEV), force_simt=False)
    sd = {k: v.to(torch.bfloat16).float() for k, v in g["state"].items()}
    xr, er = O.mgn_layer(sd, "", x.float(), ea.float(), ei, g["kwargs"]["aggregation"])
    # tensor-core path vs fp32-math path on identical bf16 inputs / weights
    assert rel_l2(xu.float(), xs.float()) < 1e-2, rel_l2(xu.float(), xs.float())
    assert rel_l2(eu.float(), es.float()) < 1e-2 if e else True
    assert rel_l2(xu.float(), xr) < 1e-2 and (e == 0 or rel_l2(eu.float(), er) < 1e-2)


def test_umma_forward_is_deterministic():
    import aero_gnn_b200.models as M
    g = load_golden("layer_sum_L2_add")
    layer = M.MeshGraphNetLayer(128, 128, 128, **g["kwargs"])
    layer.load_state_dict(g["state"])
    layer = layer.to(DEV).to(torch.bfloat16)
    gen = torch.Generator().manual_seed(11)
    n, e = 20000, 120000
    x = torch.randn(n, 128, generator=gen).to(DEV, torch.bfloat16)
    ea = torch.randn(e, 128, generator=gen).to(DEV, torch.bfloat16)
    ei = torch.randint(0, n, (2, e), generator=gen).to(DEV)
    a = _run_layer(layer, x, ea, ei, False)
    b = _run_layer(layer, x, ea, ei, False)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def _grads(layer, x, e, ei, probe, force_simt):
    old = os.environ.get("AERO_FORCE_SIMT")
    os.environ["AERO_FORCE_SIMT"] = "1" if force_simt else "0"
    try:
        for p in layer.parameters():
            p.grad = None
        xg, eg = x.clone().requires_grad_(True), e.clone().requires_grad_(True)
        xo, eo = layer(xg, eg, ei)
        (torch.cat([xo, eo], 0).float() * probe).sum().backward()
        return xg.grad.float(), eg.grad.float(), {n: p.grad.float().clone() for n, p in layer.named_parameters()}
    finally:
        if old is None:
            os.environ.pop("AERO_FORCE_SIMT", None)
        else:
            os.environ["AERO_FORCE_SIMT"] = old


@pytest.mark.parametrize("name", ["layer_sum_L2_add", "layer_cat_L1_mean"])
@pytest.mark.parametrize("n,e", [(37, 301), (300, 2111), (10, 700), (5000, 29600)])
def test_umma_backward_matches_simt_and_oracle(name, n, e):
    import aero_gnn_b200.models as M
    from aero_gnn_b200 import lib as L
    assert L.load().aero_has_umma_bwd() == 1
    g = load_golden(name)
    layer = M.MeshGraphNetLayer(128, 128, 128, **g["kwargs"])
    layer.load_state_dict(g["state"])
    layer = layer.to(DEV).to(torch.bfloat16)
    gen = torch.Generator().manual_seed(n + e)
    x = torch.randn(n, 128, generator=gen).to(torch.bfloat16)
    ea = torch.randn(e, 128, generator=gen).to(torch.bfloat16)
    ei = torch.randint(0, n, (2, e), generator=gen)
    probe = torch.randn(n + e, 128, generator=gen)
    gs = _grads(layer, x.to(DEV), ea.to(DEV), ei.to(DEV), probe.to(DEV), True)
    gu = _grads(layer, x.to(DEV), ea.to(DEV), ei.to(DEV), probe.to(DEV), False)
    # Truth: fp32 oracle autograd on the bf16-held parameters / inputs.  Yardstick: the reference's own bf16 mode
    # (train.py:30-33 = pure bf16 tensors and autograd), restated by running the oracle in bf16 on the CPU.  The
    # tensor-core path must be at least as close to the fp32 truth as that (within 2x), or within 1e-2.
    def oracle(dt):
        sd = {k: v.to(torch.bfloat16).to(dt).requires_grad_(True) for k, v in g["state"].items()}
        xr, er = x.to(dt).requires_grad_(True), ea.to(dt).requires_grad_(True)
        xo, eo = O.mgn_layer(sd, "", xr, er, ei, g["kwargs"]["aggregation"])
        names = list(sd)
        gr = torch.autograd.grad((torch.cat([xo, eo], 0).float() * probe).sum(), [xr, er] + [sd[k] for k in names])
        return names, [t.float() for t in gr]
    names, ref = oracle(torch.float32)
    _, ref16 = oracle(torch.bfloat16)

    # With a few tens of node rows (e.g. 10 rows of degree 70) a node-block gradient is a sum over very few rows and a
    # handful of ReLU sign flips of near-zero pre-activations moves it by several percent in ANY bf16 pipeline
    # (the CUDA-core fp32-math path on bf16 storage shows the same 8-10 %; the fp32 path is exact to 4e-7 on this
    # shape, scripts/diag_fp32_grads.py), so the floor of the allowance is wider for that stress shape.
    floor = 0.12 if n <= 40 else 2e-2

    def ok(mine, truth, yard, what):
        err, bar = rel_l2(mine, truth), max(floor, 2.0 * rel_l2(yard, truth))
        assert err <= bar, (what, err, bar)
        return err / bar
    worst = max(ok(gu[0], ref[0], ref16[0], "g_x"), ok(gu[1], ref[1], ref16[1], "g_e"))
    for k, gr, g16 in zip(names, ref[2:], ref16[2:]):
        worst = max(worst, ok(gu[2][k], gr, g16, k))
    # and it must stay close to the fp32-math kernels on the same bf16 data
    assert rel_l2(gu[0], gs[0]) < 6e-2 and rel_l2(gu[1], gs[1]) < 6e-2
    print("umma bwd worst error / allowance:", worst)


def test_umma_backward_is_deterministic():
    import aero_gnn_b200.models as M
    g = load_golden("layer_sum_L2_add")
    layer = M.MeshGraphNetLayer(128, 128, 128, **g["kwargs"])
    layer.load_state_dict(g["state"])
    layer = layer.to(DEV).to(torch.bfloat16)
    gen = torch.Generator().manual_seed(3)
    n, e = 30000, 200000
    x = torch.randn(n, 128, generator=gen).to(DEV, torch.bfloat16)
    ea = torch.randn(e, 128, generator=gen).to(DEV, torch.bfloat16)
    ei = torch.randint(0, n, (2, e), generator=gen).to(DEV)
    probe = torch.randn(n + e, 128, generator=gen).to(DEV)
    a = _grads(layer, x, ea, ei, probe, False)
    b = _grads(layer, x, ea, ei, probe, False)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    for k in a[2]:
        assert torch.equal(a[2][k], b[2][k]), k


@pytest.mark.parametrize("nh,agg", [(2, "mean"), (1, "add")])
def test_kept_h0_equals_recompute_bit_for_bit(monkeypatch, nh, agg):
    """The backward that reads the h_0 rows kept by the forward gives the same bits as the one that recomputes layer 0
    (aero_block_desc.h0; AERO_KEEP_H0=0 selects the recompute) when both run the first-generation kernel
    (AERO_BWD_V1=1); the TMA-fed kernel that kept h_0 rows normally select sums the residual add in fp32 instead of
    through a bf16 tile, so against the recompute it agrees to bf16 rounding, not to the bit."""
    import aero_gnn_b200.models as M
    from aero_gnn_b200 import ops
    from aero_gnn_b200.meshes import wing_surface_mesh
    from aero_gnn_b200.models._common import run_layers
    dev = torch.device("cuda", 0)
    mesh = wing_surface_mesh(37, 23)
    torch.manual_seed(9)
    net = M.MeshGraphNet(6, 4, 5, processor_size=2, num_hidden_layers_node_processor=nh,
                         num_hidden_layers_edge_processor=nh, aggregation=agg, do_concat_trick=(nh == 2)
                         ).to(dev).to(torch.bfloat16)
    plan = ops.PLAN_CACHE.get(mesh.edge_index.to(dev), mesh.num_nodes)
    g = torch.Generator().manual_seed(2)
    x0 = torch.randn(mesh.num_nodes, 128, generator=g).to(dev, torch.bfloat16).requires_grad_(True)
    e0 = torch.randn(mesh.num_edges, 128, generator=g).to(dev, torch.bfloat16).requires_grad_(True)
    gx = torch.randn(mesh.num_nodes, 128, generator=g).to(dev, torch.bfloat16)

    def run(keep):
        monkeypatch.setenv("AERO_KEEP_H0", "1" if keep else "0")
        for p in net.layers.parameters():
            p.grad = None
        x0.grad = e0.grad = None
        x, e = run_layers(net.layers, plan, x0, e0)
        torch.autograd.backward([x], [gx])
        return [x.detach().clone(), x0.grad.clone(), e0.grad.clone()] + [p.grad.clone() for p in net.layers.parameters()]

    monkeypatch.setenv("AERO_BWD_V1", "1")
    a, b = run(True), run(False)
    assert ops.keeps_h0(1, 1, 1, 1) is False       # env still "0" here
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    monkeypatch.setenv("AERO_BWD_V1", "0")
    c = run(True)                                  # TMA-fed kernel on the kept rows
    errs = [rel_l2(u.float(), v.float()) for u, v in zip(c, b)]
    print("TMA-fed kernel vs recompute, relative L2 per tensor:", ["%.2e" % e for e in errs])
    assert max(errs) < 2e-2, errs                  # bf16 rounding of one differently-summed value, amplified by 2 layers


def test_tmem_layout_probes():
    """Hardware facts the next kernel design relies on (aero_umma_probe): an M = 64 accumulator
    occupies lanes 32*(r/16) + r%16 (+16 with lane offset 16, so two of them share one column range), and a GEMM whose
    A operand was written to tensor memory with tcgen05.st reproduces a * b^T."""
    from aero_gnn_b200 import lib
    L = lib.load_probe()                       # libaero_probe.so: the probes are not in the product library
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    a = torch.randn(128, 128, generator=g).to(dev, torch.bfloat16)
    b = torch.randn(128, 128, generator=g).to(dev, torch.bfloat16)
    ref = a.float() @ b.float().t()
    st = torch.cuda.current_stream().cuda_stream
    for mode, off in ((0, 0), (1, 16)):
        c = torch.full((128, 128), float("nan"), device=dev)
        assert L.aero_umma_probe(a.data_ptr(), b.data_ptr(), c.data_ptr(), mode, st) == 0
        torch.cuda.synchronize()
        want = torch.zeros(128, 128, device=dev)
        for r in range(64):
            want[32 * (r // 16) + r % 16 + off] = ref[r]
        assert float((c - want).abs().max()) < 1e-3 * float(ref.abs().max())
    c = torch.empty(128, 128, device=dev)
    assert L.aero_umma_probe(a.data_ptr(), b.data_ptr(), c.data_ptr(), 2, st) == 0
    torch.cuda.synchronize()
    assert float((c - ref).abs().max()) < 1e-3 * float(ref.abs().max())


@pytest.mark.parametrize("name", ["layer_sum_L2_add", "layer_cat_L1_mean"])
@pytest.mark.parametrize("n,e", [(300, 2111), (129, 1), (7000, 41237)])
def test_tma_backward_matches_first_generation_kernel(name, n, e):
    """The TMA-fed backward kernel (block_umma_bwd2.cu: tensor-map loads / stores, rotating tile roles, mma.sync
    column sums, packed ReLU masks) against the first-generation kernel (AERO_BWD_V1=1) on identical inputs: the
    GEMM chains are the same instructions on the same bytes, so data gradients agree to bf16 rounding of the one
    value that is summed differently (the residual add) and weight gradients to fp32 summation order."""
    import aero_gnn_b200.models as M
    g = load_golden(name)
    layer = M.MeshGraphNetLayer(128, 128, 128, **g["kwargs"])
    layer.load_state_dict(g["state"])
    layer = layer.to(DEV).to(torch.bfloat16)
    gen = torch.Generator().manual_seed(7 * n + e)
    x = torch.randn(n, 128, generator=gen).to(DEV, torch.bfloat16)
    ea = torch.randn(e, 128, generator=gen).to(DEV, torch.bfloat16)
    ei = torch.randint(0, n, (2, e), generator=gen).to(DEV)
    probe = torch.randn(n + e, 128, generator=gen).to(DEV)
    old = os.environ.get("AERO_BWD_V1")
    try:
        os.environ["AERO_BWD_V1"] = "1"
        g1 = _grads(layer, x, ea, ei, probe, False)
        os.environ["AERO_BWD_V1"] = "0"
        g2 = _grads(layer, x, ea, ei, probe, False)
        g2b = _grads(layer, x, ea, ei, probe, False)
    finally:
        if old is None:
            os.environ.pop("AERO_BWD_V1", None)
        else:
            os.environ["AERO_BWD_V1"] = old
    assert torch.equal(g2[0], g2b[0]) and torch.equal(g2[1], g2b[1])        # deterministic
    assert all(torch.equal(g2[2][k], g2b[2][k]) for k in g2[2])
    assert rel_l2(g2[0], g1[0]) < 5e-3 and rel_l2(g2[1], g1[1]) < 5e-3, (rel_l2(g2[0], g1[0]), rel_l2(g2[1], g1[1]))
    for k in g1[2]:
        assert rel_l2(g2[2][k], g1[2][k]) < 5e-3, (k, rel_l2(g2[2][k], g1[2][k]))


@pytest.mark.parametrize("agg", ["add", "mean"])
def test_keep_all_policy_equals_recompute_bit_for_bit(monkeypatch, agg):
    """AERO_KEEP_ACTS=all (the forward also keeps H_1, H_2; the TMA-fed backward reads them instead of recomputing the
    hidden layers) gives the same bits as AERO_KEEP_ACTS=h0: the kept tiles are exactly what the recompute produces."""
    import aero_gnn_b200.models as M
    from aero_gnn_b200 import ops
    from aero_gnn_b200.meshes import wing_surface_mesh
    from aero_gnn_b200.models._common import run_layers
    dev = torch.device("cuda", 0)
    mesh = wing_surface_mesh(61, 37)             # 2257 nodes, ragged last tiles
    torch.manual_seed(4)
    net = M.MeshGraphNet(6, 4, 5, processor_size=3, num_hidden_layers_node_processor=2,
                         num_hidden_layers_edge_processor=2, aggregation=agg, do_concat_trick=True).to(dev).to(torch.bfloat16)
    plan = ops.PLAN_CACHE.get(mesh.edge_index.to(dev), mesh.num_nodes)
    g = torch.Generator().manual_seed(8)
    x0 = torch.randn(mesh.num_nodes, 128, generator=g).to(dev, torch.bfloat16).requires_grad_(True)
    e0 = torch.randn(mesh.num_edges, 128, generator=g).to(dev, torch.bfloat16).requires_grad_(True)
    gx = torch.randn(mesh.num_nodes, 128, generator=g).to(dev, torch.bfloat16)
    ge = torch.randn(mesh.num_edges, 128, generator=g).to(dev, torch.bfloat16)

    def run(mode):
        monkeypatch.setenv("AERO_KEEP_ACTS", mode)
        for p in net.layers.parameters():
            p.grad = None
        x0.grad = e0.grad = None
        x, e = run_layers(net.layers, plan, x0, e0)
        torch.autograd.backward([x, e], [gx, ge])
        return [x.detach().clone(), e.detach().clone(), x0.grad.clone(), e0.grad.clone()] + [p.grad.clone() for p in net.layers.parameters()]

    a, b = run("all"), run("h0")
    for i, (u, v) in enumerate(zip(a, b)):
        assert torch.equal(u, v), i
