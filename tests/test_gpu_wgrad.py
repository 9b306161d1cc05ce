"""aero_wgrad (csrc/wgrad.cu: TMA producer warp, tcgen05 MMA-issuer warp, receiver-sum consumer warps) against
torch.mm / aero_segment_reduce on the same bf16 rows."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("rows", [1, 127, 128, 129, 1000, 70001])
@pytest.mark.parametrize("a", [1, 2])
def test_wgrad_matches_library_gemm(rows, a):
    from aero_gnn_b200 import ops
    g = torch.Generator().manual_seed(rows + a)
    wide = torch.randn(rows, 128 * a + 128, generator=g).to(DEV, torch.bfloat16)
    A = wide[:, : 128 * a]                      # a column block of a wider matrix (row stride 128 a + 128)
    B = torch.randn(rows, 128, generator=g).to(DEV, torch.bfloat16)
    out = torch.full((128 * a, 128), float("nan"), device=DEV)
    ops.wgrad(A, B, out)
    ref = A.double().t() @ B.double()
    err = float((out.double() - ref).abs().max() / ref.abs().max().clamp(min=1e-30))
    assert err < 2e-6, err                      # exact bf16 products, fp32 accumulation over the rows
    out2 = torch.empty_like(out)
    ops.wgrad(A, B, out2)
    assert torch.equal(out, out2)               # deterministic


@pytest.mark.parametrize("n,e", [(37, 301), (300, 2111), (10, 700), (129, 1), (5000, 29600), (2000, 5)])
def test_wgrad_receiver_sums_equal_segment_reduce(n, e):
    """The receiver sums taken from the shared-memory tiles equal aero_segment_reduce: bit for bit for runs inside one
    128-row tile (same order, fp32 accumulation); runs that straddle tiles are summed per tile first, so they may
    differ by one bf16 rounding (2^-8 relative).  Receivers without any edge are written as zeros."""
    from aero_gnn_b200 import ops
    g = torch.Generator().manual_seed(n * 7 + e)
    ei = torch.randint(0, n, (2, e), generator=g).to(DEV)
    plan = ops.build_graph_plan(ei, n)
    G = torch.randn(e, 128, generator=g).to(DEV, torch.bfloat16)
    X = torch.randn(e, 128, generator=g).to(DEV, torch.bfloat16)
    out = torch.empty(128, 128, device=DEV)
    wide = torch.full((n, 256), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.wgrad(G, X, out, seg=(plan.dst, plan.rowptr, n, wide[:, 128:]))
    ref = ops.segment_reduce(G, plan.rowptr, None, n)
    got = wide[:, 128:]
    rp = plan.rowptr.long()
    inside = ((rp[:-1] // 128) == ((rp[1:] - 1).clamp(min=0) // 128)) | (rp[1:] == rp[:-1])
    assert torch.equal(got[inside], ref[inside])
    d = (got.float() - ref.float()).abs()
    assert bool((d <= ref.float().abs() * 2.0 ** -7 + 1e-4).all()), float(d.max())
    assert bool(torch.isnan(wide[:, :128].float()).all())          # the other column block is untouched
    refw = G.double().t() @ X.double()
    assert float((out.double() - refw).abs().max() / refw.abs().max()) < 2e-6


def test_wgrad_zero_rows_writes_zeros():
    """No rows (a coarse level without edges): dW = 0 and every receiver sum = 0, no kernel reads a null pointer."""
    from aero_gnn_b200 import ops
    n = 77
    plan = ops.build_graph_plan(torch.zeros(2, 0, dtype=torch.long, device=DEV), n)
    G = torch.empty(0, 128, device=DEV, dtype=torch.bfloat16)
    out = torch.full((128, 128), float("nan"), device=DEV)
    wide = torch.full((n, 256), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.wgrad(G, G.clone(), out, seg=(plan.dst, plan.rowptr, n, wide[:, 128:]))
    assert float(out.abs().max()) == 0.0 and float(wide[:, 128:].float().abs().max()) == 0.0
