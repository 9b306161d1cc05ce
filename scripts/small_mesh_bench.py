"""C2 (the reference's own regime): a batch of 8 airfoil meshes of 5,000 nodes (N = 40,000, E = 236,800), MGN-15, bf16,
whole training step through the public model API (forward + MSE loss + backward): eager launches vs one CUDA-graph
replay (aero_gnn_b200.graphs.GraphedStep).  Prints ms/step and edges/s for both."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aero_gnn_b200.models as M
from aero_gnn_b200 import ops
from aero_gnn_b200.graphs import GraphedStep
from aero_gnn_b200.meshes import airfoil_o_mesh, batch_meshes
import bench

dev = torch.device("cuda", 0)
mesh = batch_meshes([airfoil_o_mesh(100, 50, seed=s) for s in range(8)])
torch.manual_seed(0)
kw = dict(bench.CFG)
net = M.MeshGraphNet(mesh.node_attr.size(1), mesh.edge_attr.size(1), mesh.target.size(1), **kw).to(dev).to(torch.bfloat16)
na, ea = mesh.node_attr.to(dev, torch.bfloat16), mesh.edge_attr.to(dev, torch.bfloat16)
ei, tg = mesh.edge_index.to(dev), mesh.target.to(dev)
lossf = torch.nn.MSELoss()
out = {}


def step():
    net.zero_grad(set_to_none=True)
    pred = net(na, ea, ei)
    loss = lossf(pred.float(), tg)
    loss.backward()
    out["loss"] = loss.detach()


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


E = mesh.num_edges
ms_eager = timeit(step)
l_eager = float(out["loss"])
g = GraphedStep(step)
ms_graph = timeit(g)
l_graph = float(out["loss"])
print(f"C2 batch: N={mesh.num_nodes} E={E}  eager {ms_eager:.2f} ms/step ({E / ms_eager / 1e3:.2f} M edges/s)  "
      f"graph replay {ms_graph:.2f} ms/step ({E / ms_graph / 1e3:.2f} M edges/s)  launches/step {g.launches}  "
      f"loss eager {l_eager:.6f} graph {l_graph:.6f}")
