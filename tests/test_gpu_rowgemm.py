"""aero_row_gemm (csrc/rowgemm.cu: TMA producer warp, tcgen05 MMA-issuer warp, 8 epilogue warps, TMA stores) against
fp64 references: the sum-trick pre-projection P = x W^T + b and the back-projection g_x = [g_Ps | g_Pd | g_h0n] W + G_x."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _bf(t):
    return t.to(DEV, torch.bfloat16)


@pytest.mark.parametrize("rows", [1, 127, 128, 129, 1000, 40001])
@pytest.mark.parametrize("nb", [1, 2, 3])
def test_projection_matches_fp64(rows, nb):
    from aero_gnn_b200 import ops
    g = torch.Generator().manual_seed(rows * 3 + nb)
    x, W, b = _bf(torch.randn(rows, 128, generator=g)), _bf(torch.randn(128 * nb, 128, generator=g) / 11), _bf(torch.randn(128 * nb, generator=g))
    out = ops.row_gemm([x], W, w_mn=False, nb=nb, bias=b)
    ref = x.double() @ W.double().t() + b.double()
    assert out.shape == (rows, 128 * nb)
    err = (out.double() - ref).abs()
    assert bool((err <= ref.abs() * 2.0 ** -8 + 1e-6).all()), float(err.max())      # one bf16 rounding of an fp32 sum
    assert torch.equal(out, ops.row_gemm([x], W, w_mn=False, nb=nb, bias=b))          # deterministic


@pytest.mark.parametrize("rows", [1, 130, 5000, 33333])
@pytest.mark.parametrize("na", [1, 2, 3])
def test_back_projection_matches_fp64(rows, na):
    """K-blocks that are column blocks of wider matrices, the addend added in fp32 before the single rounding, the
    result written into a column block of a wider matrix."""
    from aero_gnn_b200 import ops
    g = torch.Generator().manual_seed(rows * 5 + na)
    wide = _bf(torch.randn(rows, 256, generator=g))
    third = _bf(torch.randn(rows, 128, generator=g))
    blocks = [wide[:, :128], wide[:, 128:], third][:na]
    W = _bf(torch.randn(128 * na, 128, generator=g) / 13)
    G = _bf(torch.randn(rows, 128, generator=g))
    dest = torch.full((rows, 384), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.row_gemm(blocks, W, w_mn=True, add=G, out=dest[:, 128:256])
    ref = torch.cat([b.double() for b in blocks], 1) @ W.double() + G.double()
    err = (dest[:, 128:256].double() - ref).abs()
    assert bool((err <= ref.abs() * 2.0 ** -8 + 1e-5).all()), float(err.max())
    assert bool(torch.isnan(dest[:, :128].float()).all()) and bool(torch.isnan(dest[:, 256:].float()).all())
    plain = ops.row_gemm(blocks, W, w_mn=True)                                         # no addend, fresh output
    ref2 = torch.cat([b.double() for b in blocks], 1) @ W.double()
    assert bool(((plain.double() - ref2).abs() <= ref2.abs() * 2.0 ** -8 + 1e-5).all())


def test_row_gemm_zero_rows_and_bad_shapes():
    from aero_gnn_b200 import ops
    W = _bf(torch.randn(384, 128))
    assert ops.row_gemm([torch.empty(0, 128, device=DEV, dtype=torch.bfloat16)], W, w_mn=False, nb=3).shape == (0, 384)
    with pytest.raises(RuntimeError):
        ops.row_gemm([_bf(torch.randn(4, 128))], W, w_mn=False, nb=2)      # W has 3 tiles, nb = 2 asks for 2
    with pytest.raises(RuntimeError):
        ops.row_gemm([torch.randn(4, 128, device=DEV)], W[:128].contiguous(), w_mn=False)   # fp32 rows
