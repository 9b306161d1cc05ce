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
