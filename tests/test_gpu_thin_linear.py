"""aero_thin_linear_{fwd,bwd} (the encoders' first Linear on raw features) against torch.nn.functional.linear."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("K", [1, 4, 6, 9, 16])
@pytest.mark.parametrize("rows", [1, 33, 5000])
def test_thin_linear_forward_backward(dtype, tol, K, rows):
    from aero_gnn_b200 import ops
    g = torch.Generator().manual_seed(K * 1000 + rows)
    x = torch.randn(rows, K, generator=g).to(DEV, dtype)
    W = (torch.randn(128, K, generator=g) / 3).to(DEV, dtype).requires_grad_(True)
    b = torch.randn(128, generator=g).to(DEV, dtype).requires_grad_(True)
    G = torch.randn(rows, 128, generator=g).to(DEV, dtype)
    out = ops.ThinLinearFn.apply(x, W, b)
    out.backward(G)
    gW, gb = W.grad.clone(), b.grad.clone()
    ref = F.linear(x.double(), W.detach().double(), b.detach().double())
    assert float((out.double() - ref).abs().max()) <= tol * max(1.0, float(ref.abs().max()))
    refW = G.double().t() @ x.double()
    refb = G.double().sum(0)
    assert float((gW.double() - refW).abs().max()) <= tol * max(1.0, float(refW.abs().max()))
    assert float((gb.double() - refb).abs().max()) <= tol * max(1.0, float(refb.abs().max()))
    W.grad = b.grad = None
    ops.ThinLinearFn.apply(x, W, b).backward(G)
    assert torch.equal(W.grad, gW) and torch.equal(b.grad, gb)          # deterministic


def test_thin_linear_input_gradient_and_strided_rows():
    from aero_gnn_b200 import ops
    wide = torch.randn(100, 12, device=DEV)
    x = wide[:, 2:8].detach().requires_grad_(True)                       # row stride 12, 6 features
    W = torch.randn(128, 6, device=DEV, requires_grad=True)
    b = torch.randn(128, device=DEV, requires_grad=True)
    out = ops.ThinLinearFn.apply(x, W, b)
    ref = F.linear(x.detach(), W.detach(), b.detach())
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5)
    out.square().sum().backward()
    assert torch.allclose(x.grad, (2 * ref) @ W.detach(), rtol=1e-4, atol=1e-4)
