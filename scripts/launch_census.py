"""Which ops launch the kernels of one processor step: per autograd/aten op, the number of CUDA kernels and their time
(small mesh, so the counts -- not the times -- are what matters).  usage: launch_census.py [layers]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import aero_gnn_b200.models as M
from aero_gnn_b200 import ops
from aero_gnn_b200.meshes import wing_surface_mesh
from aero_gnn_b200.models._common import run_layers
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = "cuda:0"
mesh = wing_surface_mesh(200, 100)
kw = dict(processor_size=layers, num_hidden_layers_node_processor=2, num_hidden_layers_edge_processor=2,
          num_hidden_layers_node_encoder=2, num_hidden_layers_edge_encoder=2, num_hidden_layers_decoder=2,
          aggregation="add", do_concat_trick=True)
net = M.MeshGraphNet(6, 4, 5, **kw).to(dev).to(torch.bfloat16)
plan = ops.PLAN_CACHE.get(mesh.edge_index.to(dev), mesh.num_nodes)
g = torch.Generator().manual_seed(1)
x0 = torch.randn(mesh.num_nodes, 128, generator=g).to(dev, torch.bfloat16).requires_grad_(True)
e0 = torch.randn(mesh.num_edges, 128, generator=g).to(dev, torch.bfloat16).requires_grad_(True)
gx = torch.ones(mesh.num_nodes, 128, device=dev, dtype=torch.bfloat16)


def step():
    for p in net.layers.parameters():
        p.grad = None
    x0.grad = e0.grad = None
    x, e = run_layers(net.layers, plan, x0, e0)
    torch.autograd.backward([x], [gx])


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
evs = prof.events()
kern = [e for e in evs if e.device_type == torch.autograd.DeviceType.CUDA]
print("kernels in one step:", len(kern), "layers:", layers, "=> per layer", len(kern) / layers)
by_name = collections.Counter(k.name[:90] for k in kern)
for n, c in by_name.most_common(40):
    print(f"{c:5d}  {n}")
# attribute kernels to the innermost CPU op that launched them (correlation through time containment)
cpu = [e for e in evs if e.device_type == torch.autograd.DeviceType.CPU]
print("---- CPU ops that own kernels (count of kernels launched inside, by op name) ----")
own = collections.Counter()
for e in cpu:
    if e.kernels:
        own[e.name[:70]] += len(e.kernels)
for n, c in own.most_common(40):
    print(f"{c:5d}  {n}")
