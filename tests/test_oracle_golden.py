"""Pin the CPU oracle (oracle/mgn_oracle.py) against fixtures produced by the unmodified reference
(oracle/gen_golden.py).  fp32 tolerance: 1e-5 relative (north_star), integers exact."""
import numpy as np
import torch

from conftest import load_golden, rel_err
from oracle import mgn_oracle as O

TOL = 1e-5


def _autograd(out, probe, wrt):
    return torch.autograd.grad((out * probe).sum(), wrt, allow_unused=True)


def test_mlp():
    g = load_golden("mlp")
    assert rel_err(O.mlp(g["state"], "", g["x"]), g["out"]) < TOL


def test_standalone_blocks():
    g = load_golden("blocks")
    assert rel_err(O.edge_block(g["state_eb"], "", g["e"], g["x"], g["edge_index"]), g["out_eb"]) < TOL
    assert rel_err(O.edge_block(g["state_es"], "", g["e"], g["x"], g["edge_index"]), g["out_es"]) < TOL
    assert rel_err(O.node_block(g["state_nb"], "", g["x"], g["e"], g["edge_index"], "mean"), g["out_nb"]) < TOL


def test_layers_forward_backward():
    for name in ("layer_sum_L2_add", "layer_cat_L1_mean"):
        g = load_golden(name)
        sd = {k: v.clone().requires_grad_(True) for k, v in g["state"].items()}
        x = g["x"].clone().requires_grad_(True)
        e = g["e"].clone().requires_grad_(True)
        xo, eo = O.mgn_layer(sd, "", x, e, g["edge_index"], g["kwargs"]["aggregation"])
        assert rel_err(xo, g["x_out"]) < TOL and rel_err(eo, g["e_out"]) < TOL
        names = list(sd)
        grads = _autograd(torch.cat([xo, eo], 0), g["probe"], [x, e] + [sd[n] for n in names])
        assert rel_err(grads[0], g["g_x"]) < 1e-4 and rel_err(grads[1], g["g_e"]) < 1e-4
        for n, gr in zip(names, grads[2:]):
            assert rel_err(gr, g["g_params"][n]) < 1e-4, n


def test_mgn_model():
    g = load_golden("mgn")
    out = O.mgn_forward(g["state"], g["node_attr"], g["edge_attr"], g["edge_index"], "add")
    assert rel_err(out, g["out"]) < TOL


def test_fourier_and_pool_models():
    g = load_golden("fouriermgn")
    assert torch.equal(O.fourier_embedding(g["node_attr"]), g["emb"])
    assert rel_err(O.fourier_mgn_forward(g["state"], g["node_attr"], g["edge_attr"], g["edge_index"]), g["out"]) < TOL
    g = load_golden("poolmgn")
    out = O.pool_mgn_forward(g["state"], g["node_attr"], g["edge_attr"], g["edge_index"], g["batch"], "mean")
    assert rel_err(out, g["out"]) < TOL


def test_bsms_indices_exact():
    g = load_golden("bsms")
    f2c, cb = O.stride_pool_indices(g["batch"].numpy(), g["pos"][:, 0].numpy(), 2)
    assert np.array_equal(f2c, g["l1_f2c"].numpy()) and np.array_equal(cb, g["l1_cb"].numpy())
    cei, _ = O.coarsen_edge_indices(g["edge_index"].numpy(), f2c, cb.shape[0])
    assert np.array_equal(cei, g["l1_cei"].numpy())
    # second level is driven by the mean-pooled positions of the first
    f2c2, cb2 = O.stride_pool_indices(g["l1_cb"].numpy(), g["l1_cpos"][:, 0].numpy(), 2)
    assert np.array_equal(f2c2, g["l2_f2c"].numpy()) and np.array_equal(cb2, g["l2_cb"].numpy())
    cei2, _ = O.coarsen_edge_indices(g["l1_cei"].numpy(), f2c2, cb2.shape[0])
    assert np.array_equal(cei2, g["l2_cei"].numpy())
    # no positions: index order
    f2c0, cb0 = O.stride_pool_indices(g["batch"].numpy(), None, 2)
    assert np.array_equal(f2c0, g["nopos_f2c"].numpy()) and np.array_equal(cb0, g["nopos_cb"].numpy())
    s3 = load_golden("bsms_stride3")
    f2c3, cb3 = O.stride_pool_indices(s3["batch"].numpy(), s3["pos"][:, 0].numpy(), 3)
    assert np.array_equal(f2c3, s3["f2c"].numpy()) and np.array_equal(cb3, s3["cb"].numpy())
    cei3, _ = O.coarsen_edge_indices(s3["edge_index"].numpy(), f2c3, cb3.shape[0])
    assert np.array_equal(cei3, s3["cei"].numpy())


def test_bsms_downsample_and_model():
    g = load_golden("bsms")
    sd = g["state"]
    xh = O.mlp(sd, "node_encoder.", g["node_attr"])
    eh = O.mlp(sd, "edge_encoder.", g["edge_attr"])
    cx, ce, cei, cb, cpos, f2c = O.downsample(xh, eh, g["edge_index"], g["batch"], g["pos"], 2)
    assert torch.equal(cpos, g["l1_cpos"])            # stride-2 means are exactly rounded
    assert rel_err(cx, g["l1_cx"]) < TOL and rel_err(ce, g["l1_ce"]) < TOL
    out = O.bsms_forward(sd, g["node_attr"], g["edge_attr"], g["edge_index"], g["batch"], g["pos"], 2)
    assert rel_err(out, g["out"]) < TOL


def test_receiver_csr_matches_stable_sort():
    rng = np.random.default_rng(0)
    n, e = 50, 400
    ei = rng.integers(0, n, size=(2, e))
    rowptr, perm, src, dst, sptr, sperm = O.receiver_csr(ei, n)
    order = torch.sort(torch.from_numpy(ei[1]), stable=True).indices.numpy()
    assert np.array_equal(perm, order)
    assert np.all(np.diff(dst) >= 0) and rowptr[-1] == e and sptr[-1] == e
    assert np.array_equal(np.sort(sperm), np.arange(e))
    assert np.all(np.diff(src[sperm]) >= 0)
