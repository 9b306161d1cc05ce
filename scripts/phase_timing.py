"""Build with AERO_NVCC_EXTRA=-DAERO_PHASE_TIMING, then: per-phase cycles of the tcgen05 backward kernel (CTA 0)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aero_gnn_b200.models as M
from aero_gnn_b200 import ops, lib
from aero_gnn_b200.meshes import wing_surface_mesh
from aero_gnn_b200.models._common import run_layers
dev = "cuda:0"
mesh = wing_surface_mesh(1000, 1000)
kw = dict(processor_size=1, num_hidden_layers_node_processor=2, num_hidden_layers_edge_processor=2, aggregation="add", do_concat_trick=True)
net = M.MeshGraphNet(6, 4, 5, **kw).to(dev).to(torch.bfloat16)
plan = ops.PLAN_CACHE.get(mesh.edge_index.to(dev), mesh.num_nodes)
g = torch.Generator().manual_seed(1)
x0 = torch.randn(mesh.num_nodes, 128, generator=g).to(dev, torch.bfloat16).requires_grad_(True)
e0 = torch.randn(mesh.num_edges, 128, generator=g).to(dev, torch.bfloat16).requires_grad_(True)
L = lib.load()
fn = L.aero_debug_phase_read; fn.argtypes = [C.c_void_p, C.c_int]; fn.restype = C.c_int
buf = (C.c_longlong * 32)()
names = ["tile-top sync", "staging", "fwd issue+colsum", "fwd MMA wait", "fwd epilogues", "LN backward", "bwd issue+colsum", "bwd MMA wait", "bwd epilogues", "-", "bwd tcgen05.ld", "bwd mask+store", "bwd fences", "bwd barrier", "bwd pre-colsum", "(lane0) fwd weight wait", "(lane0) fwd issue", "(lane0) bwd weight wait", "(lane0) bwd issue", "(lane0) to issue point"]
for it in range(3):
    x, e = run_layers(net.layers, plan, x0, e0)
    fn(buf, 1)       # reset after forward
    torch.autograd.backward([x], [torch.ones_like(x)])
    fn(buf, 1)
    v = list(buf)[:20]
    # the backward runs the node block kernel then the edge block kernel: both accumulate; report the sum
    tot = sum(v)
    print("iter", it, "total cycles (CTA0 observer, node+edge bwd kernels):", tot)
    for n, c in zip(names, v):
        print(f"   {n:22s} {c:>12d}  {100*c/max(tot,1):5.1f}%")
