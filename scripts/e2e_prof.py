"""torch.profiler breakdown of the end-to-end step (MeshGraphNet.forward + MSELoss + backward) at C5."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import aero_gnn_b200.models as M
from aero_gnn_b200.meshes import wing_surface_mesh
import bench
nu = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = "cuda:0"
mesh = wing_surface_mesh(nu, nu)
torch.manual_seed(0)
net = M.MeshGraphNet(6, 4, 5, **bench.CFG).to(dev).to(torch.bfloat16)
na, ea, ei, tg = (t.to(dev) for t in (mesh.node_attr, mesh.edge_attr, mesh.edge_index, mesh.target))
lossf = torch.nn.MSELoss()
def step():
    net.zero_grad(set_to_none=True)
    pred = net(na.to(torch.bfloat16), ea.to(torch.bfloat16), ei)
    from aero_gnn_b200.train_tail import mse_loss
    loss = mse_loss(pred, tg)
    loss.backward()
    return float(loss.item())
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=60))
