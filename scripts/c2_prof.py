import os, sys
sys.path.insert(0, "/root/repo")
import torch
from torch.profiler import profile, ProfilerActivity
import aero_gnn_b200.models as M
from aero_gnn_b200.meshes import airfoil_o_mesh, batch_meshes
import bench
dev = torch.device("cuda", 0)
mesh = batch_meshes([airfoil_o_mesh(100, 50, seed=s) for s in range(8)])
torch.manual_seed(0)
net = M.MeshGraphNet(mesh.node_attr.size(1), mesh.edge_attr.size(1), mesh.target.size(1), **bench.CFG).to(dev).to(torch.bfloat16)
na, ea = mesh.node_attr.to(dev, torch.bfloat16), mesh.edge_attr.to(dev, torch.bfloat16)
ei, tg = mesh.edge_index.to(dev), mesh.target.to(dev)
lossf = torch.nn.MSELoss()
def step():
    net.zero_grad(set_to_none=True)
    loss = lossf(net(na, ea, ei).float(), tg)
    loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))
