"""Training-step tail (aero_gnn_b200/train_tail.py: fused MSE + multi-tensor Adam) against torch's own
nn.MSELoss / torch.optim.Adam -- the calls the reference's loop makes (utils.py:191-195, train.py:207-211, :222)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("rows,cols,ld", [(1000, 4, 4), (40000, 4, 128), (1, 5, 8), (0, 3, 3)])
def test_mse_loss_and_gradient_match_torch(dtype, tol, rows, cols, ld):
    from aero_gnn_b200.train_tail import mse_loss
    g = torch.Generator().manual_seed(rows + cols)
    base = torch.randn(rows, ld, generator=g).to(DEV, dtype)
    pred = base[:, :cols].detach().requires_grad_(True) if ld == cols else None
    if pred is None:   # strided view of a wider matrix (the decoder's 128-wide tile)
        wide = base.detach().requires_grad_(True)
        pred = wide[:, :cols]
    tgt = torch.randn(rows, cols, generator=g).to(DEV)
    loss = mse_loss(pred, tgt)
    assert loss.dim() == 0 and loss.dtype == torch.float32 and loss.is_cuda
    if rows == 0:
        return
    loss.backward()
    ref_in = base[:, :cols].detach().float().requires_grad_(True)
    ref = torch.nn.MSELoss()(ref_in, tgt)
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref)))
    got = (wide.grad[:, :cols] if ld != cols else pred.grad).float()
    assert float((got - ref_in.grad).abs().max()) <= tol * float(ref_in.grad.abs().max()) + 1e-12
    if ld != cols:
        assert float(wide.grad[:, cols:].abs().max()) == 0.0


@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_fused_adam_matches_torch_adam_fp32(wd):
    from aero_gnn_b200.train_tail import FusedAdam
    g = torch.Generator().manual_seed(5)
    shapes = [(128, 128), (128,), (384, 128), (7,), (1,), (33, 5)]
    mine = [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in mine]
    opt = FusedAdam(mine, lr=1e-3, weight_decay=wd)
    topt = torch.optim.Adam(ref, lr=1e-3, weight_decay=wd)
    for it in range(5):
        for a, b in zip(mine, ref):
            gr = torch.randn(a.shape, generator=g).to(DEV)
            if it == 2 and a.numel() == 7:
                a.grad = b.grad = None          # a parameter without a gradient is skipped, like torch
                continue
            a.grad, b.grad = gr.clone(), gr.clone()
        opt.step()
        topt.step()
    for a, b in zip(mine, ref):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), float((a - b).abs().max())


def test_fused_adam_bf16_params_and_master_weights():
    from aero_gnn_b200.train_tail import FusedAdam
    g = torch.Generator().manual_seed(6)
    w0 = torch.randn(256, 128, generator=g)
    grads = [torch.randn(256, 128, generator=g) * 0.1 for _ in range(20)]
    ref = torch.nn.Parameter(w0.to(DEV))
    topt = torch.optim.Adam([ref], lr=1e-2)
    plain = torch.nn.Parameter(w0.to(DEV, torch.bfloat16))
    mast = torch.nn.Parameter(w0.to(DEV, torch.bfloat16))
    o1, o2 = FusedAdam([plain], lr=1e-2), FusedAdam([mast], lr=1e-2, master_weights=True)
    for gr in grads:
        ref.grad = gr.to(DEV)
        plain.grad = gr.to(DEV, torch.bfloat16)
        mast.grad = gr.to(DEV, torch.bfloat16)
        topt.step(); o1.step(); o2.step()
    e_plain = float((plain.float() - ref).norm() / ref.norm())
    e_mast = float((mast.float() - ref).norm() / ref.norm())
    assert e_mast < 5e-3 and e_plain < 3e-2 and e_mast <= e_plain, (e_plain, e_mast)   # master copy: no accumulated rounding


def test_training_step_tail_end_to_end_small_mgn():
    """model fwd -> fused MSE -> backward -> FusedAdam, 3 steps, against the same loop with torch's loss / Adam."""
    import copy
    import aero_gnn_b200.models as M
    from aero_gnn_b200.meshes import airfoil_o_mesh
    from aero_gnn_b200.train_tail import FusedAdam, mse_loss
    mesh = airfoil_o_mesh(30, 12, seed=0)
    torch.manual_seed(0)
    kw = dict(processor_size=2, num_hidden_layers_node_processor=2, num_hidden_layers_edge_processor=2,
              num_hidden_layers_node_encoder=2, num_hidden_layers_edge_encoder=2, num_hidden_layers_decoder=2,
              aggregation="add", do_concat_trick=True)
    a = M.MeshGraphNet(6, 3, 4, **kw).to(DEV)
    b = copy.deepcopy(a)
    na, ea, ei, tg = mesh.node_attr.to(DEV), mesh.edge_attr.to(DEV), mesh.edge_index.to(DEV), mesh.target.to(DEV)
    oa, ob = FusedAdam(a.parameters(), lr=1e-3), torch.optim.Adam(b.parameters(), lr=1e-3)
    for _ in range(3):
        la = mse_loss(a(na, ea, ei), tg)
        la.backward(); oa.step(); oa.zero_grad()
        lb = torch.nn.MSELoss()(b(na, ea, ei), tg)
        lb.backward(); ob.step(); ob.zero_grad()
        assert abs(float(la) - float(lb)) < 1e-5 * max(1.0, float(lb))
    # Adam normalises every gradient element to ~ +-lr, so an element whose gradient is at rounding level (1e-9) may
    # move by up to lr per step in either loop; the loops agree in the mean to 1e-5 and nowhere differ by more than
    # the 3 x lr such an element can travel
    for (k, p), q in zip(a.named_parameters(), b.parameters()):
        d = (p - q).abs()
        assert float(d.mean()) < 1e-5 and float(d.max()) <= 3.1e-3, (k, float(d.mean()), float(d.max()))
