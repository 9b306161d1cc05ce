"""The partitioned processor on ONE rank (no halo, no process group) must reproduce the plain processor stack:
covers the partition's own forward / backward orchestration (row GEMMs, weight-gradient reductions, gradient sink)
on the GPU; the halo paths are covered by the gloo tests (CPU) and scripts/check_partition.py (torchrun, N GPUs)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_world1_partition_equals_plain_stack(dtype, tol):
    import aero_gnn_b200.models as M
    from aero_gnn_b200 import ops
    from aero_gnn_b200.meshes import wing_surface_mesh
    from aero_gnn_b200.models._common import run_layers
    from aero_gnn_b200.partition import PartitionedProcessor

    mesh = wing_surface_mesh(60, 40)
    N, E = mesh.num_nodes, mesh.num_edges
    torch.manual_seed(0)
    kw = dict(processor_size=3, num_hidden_layers_node_processor=2, num_hidden_layers_edge_processor=2,
              aggregation="add", do_concat_trick=True)
    net = M.MeshGraphNet(6, 4, 5, **kw).to(DEV).to(dtype)
    g = torch.Generator().manual_seed(7)
    xg, eg = torch.randn(N, 128, generator=g).to(dtype), torch.randn(E, 128, generator=g).to(dtype)
    probe = torch.randn(N, 128, generator=g).to(DEV)

    plan = ops.PLAN_CACHE.get(mesh.edge_index.to(DEV), N)
    x0 = xg.to(DEV).requires_grad_(True)
    e0 = eg.to(DEV)[plan.perm.long()].requires_grad_(True)
    xr, _ = run_layers(net.layers, plan, x0, e0)
    (xr.float() * probe).sum().backward()
    ref_gx, ref_ge = x0.grad.clone(), e0.grad.clone()
    ref_gw = [p.grad.clone() for p in net.layers.parameters()]
    for p in net.layers.parameters():
        p.grad = None

    pp = PartitionedProcessor(mesh.edge_index, N, 0, 1, torch.device(DEV))
    assert pp.n_own == N and pp.halo.n_halo == 0
    x1 = xg.to(DEV).requires_grad_(True)
    ids = pp.csr_edge_ids()
    e1 = eg.to(DEV)[ids].requires_grad_(True)
    xo, _ = pp.run(net.layers, x1, e1)
    (xo.float() * probe).sum().backward()

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30))
    assert rel(xo, xr) < tol
    assert rel(x1.grad, ref_gx) < 2 * tol
    assert rel(e1.grad, ref_ge[plan.inv_perm.long()[ids]]) < 2 * tol
    for p, r in zip(net.layers.parameters(), ref_gw):
        assert rel(p.grad, r) < 2 * tol
