"""Small driver for ncu: a few forward+backward processor steps (2 layers) on a mid-size wing mesh, bf16."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aero_gnn_b200.models as M
from aero_gnn_b200 import ops
from aero_gnn_b200.meshes import wing_surface_mesh
from aero_gnn_b200.models._common import run_layers
nu = int(sys.argv[1]) if len(sys.argv) > 1 else 300
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 300
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = "cuda:0"
mesh = wing_surface_mesh(nu, nv)
torch.manual_seed(0)
kw = dict(processor_size=2, num_hidden_layers_node_processor=2, num_hidden_layers_edge_processor=2,
          num_hidden_layers_node_encoder=2, num_hidden_layers_edge_encoder=2, num_hidden_layers_decoder=2,
          aggregation="add", do_concat_trick=True)
net = M.MeshGraphNet(6, 4, 5, **kw).to(dev).to(torch.bfloat16)
plan = ops.PLAN_CACHE.get(mesh.edge_index.to(dev), mesh.num_nodes)
g = torch.Generator().manual_seed(1)
x0 = torch.randn(mesh.num_nodes, 128, generator=g).to(dev, torch.bfloat16).requires_grad_(True)
e0 = torch.randn(mesh.num_edges, 128, generator=g).to(dev, torch.bfloat16).requires_grad_(True)
for _ in range(iters):
    x, e = run_layers(net.layers, plan, x0, e0)
    torch.autograd.backward([x], [torch.ones_like(x)])
torch.cuda.synchronize()
print("ok", mesh.num_nodes, mesh.num_edges)
