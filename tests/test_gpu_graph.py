"""CUDA-graph replay of a whole processor step equals the eager step bit for bit (the kernels are deterministic)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(dtype):
    import aero_gnn_b200.models as M
    from aero_gnn_b200 import ops
    from aero_gnn_b200.meshes import wing_surface_mesh
    dev = torch.device("cuda", 0)
    mesh = wing_surface_mesh(40, 30)
    torch.manual_seed(3)
    net = M.MeshGraphNet(6, 4, 5, processor_size=3, num_hidden_layers_node_processor=2,
                         num_hidden_layers_edge_processor=2, aggregation="add", do_concat_trick=True).to(dev).to(dtype)
    plan = ops.PLAN_CACHE.get(mesh.edge_index.to(dev), mesh.num_nodes)
    g = torch.Generator().manual_seed(5)
    x0 = torch.randn(mesh.num_nodes, 128, generator=g).to(dev, dtype).requires_grad_(True)
    e0 = torch.randn(mesh.num_edges, 128, generator=g).to(dev, dtype).requires_grad_(True)
    gx = torch.randn(mesh.num_nodes, 128, generator=g).to(dev, dtype)
    return net, plan, x0, e0, gx


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_graph_replay_matches_eager(dtype):
    from aero_gnn_b200 import ops
    from aero_gnn_b200.graphs import GraphedStep
    from aero_gnn_b200.models._common import run_layers
    net, plan, x0, e0, gx = _setup(dtype)
    out = {}

    def step():
        for p in net.layers.parameters():
            p.grad = None
        x0.grad = e0.grad = None
        x, e = run_layers(net.layers, plan, x0, e0)
        torch.autograd.backward([x], [gx])
        out["x"] = x.detach()       # no reference to the autograd graph survives the step
        return out["x"]

    step()
    torch.cuda.synchronize()
    ref_x = out["x"].detach().clone()
    ref_gx0, ref_ge0 = x0.grad.clone(), e0.grad.clone()
    ref_gw = [p.grad.clone() for p in net.layers.parameters()]

    l0 = ops.LaunchCounter.total
    g = GraphedStep(step)
    assert g.launches > 0
    # new inputs written in place, then replay: compare against an eager run on the same values
    with torch.no_grad():
        x0.mul_(0.5)
        e0.add_(0.25)
    l1 = ops.LaunchCounter.total
    xg = g()
    torch.cuda.synchronize()
    assert ops.LaunchCounter.total - l1 == g.launches
    got_x = xg.detach().clone()
    got_gx0, got_ge0 = x0.grad.clone(), e0.grad.clone()
    got_gw = [p.grad.clone() for p in net.layers.parameters()]
    assert not torch.equal(got_x, ref_x)            # the replay really recomputed on the new inputs

    step()                                           # eager on the modified inputs
    torch.cuda.synchronize()
    assert torch.equal(out["x"], got_x)
    assert torch.equal(x0.grad, got_gx0) and torch.equal(e0.grad, got_ge0)
    for a, b in zip(got_gw, [p.grad for p in net.layers.parameters()]):
        assert torch.equal(a, b)
    assert l0 < l1
    del ref_gx0, ref_ge0, ref_gw


def test_replays_rewrite_gradients_instead_of_accumulating():
    """GraphedStep(leaves=...) resets the warm-up's gradients before the capture, so two replays on the same inputs
    leave the same gradients (not doubled), equal to one eager step, in the static tensors of `g.grads`."""
    from aero_gnn_b200.graphs import GraphedStep
    from aero_gnn_b200.models._common import run_layers
    net, plan, x0, e0, gx = _setup(torch.bfloat16)
    params = list(net.layers.parameters())

    def step():                                   # note: does NOT clear gradients itself
        x, e = run_layers(net.layers, plan, x0, e0)
        torch.autograd.backward([x], [gx])
        return x.detach()

    g = GraphedStep(step, leaves=[x0, e0, *params])
    g()
    torch.cuda.synchronize()
    first = [t.clone() for t in g.grads]
    g()
    torch.cuda.synchronize()
    for a, b, leaf in zip(first, g.grads, [x0, e0, *params]):
        assert torch.equal(a, b)
        assert leaf.grad is b                     # the static tensor is the leaf's .grad
    for t in [x0, e0, *params]:
        t.grad = None
    step()
    torch.cuda.synchronize()
    for a, leaf in zip(first, [x0, e0, *params]):
        assert torch.equal(a, leaf.grad)
