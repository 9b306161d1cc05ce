"""torch.profiler device-time breakdown of one partitioned processor step (fwd+bwd+grad all-reduce) on rank 0:
which kernels (fused blocks, NCCL send/recv, all-reduce, library GEMMs, copies) the per-rank step time goes to.
    torchrun --nproc-per-node N scripts/scale_prof.py [nu nv]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
import aero_gnn_b200.models as M
from aero_gnn_b200.meshes import wing_surface_mesh
from aero_gnn_b200.partition import PartitionedProcessor
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nu = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
mesh = wing_surface_mesh(nu, nv)
N, E = mesh.num_nodes, mesh.num_edges
torch.manual_seed(0)
net = M.MeshGraphNet(6, 4, 5, **bench.CFG).to(dev).to(torch.bfloat16)
pp = PartitionedProcessor(mesh.edge_index, N, rank, world, dev)
g = torch.Generator().manual_seed(1234)
x0 = torch.randn(N, 128, generator=g)[pp.lo:pp.hi].to(dev, torch.bfloat16).requires_grad_(True)
e0 = torch.randn(E, 128, generator=g)[pp.edge_ids_cpu].to(dev, torch.bfloat16).requires_grad_(True)
gx = torch.ones(pp.n_own, 128, device=dev, dtype=torch.bfloat16)


def step():
    for p in net.layers.parameters():
        p.grad = None
    x0.grad = e0.grad = None
    x, e = pp.run(net.layers, x0, e0)
    torch.autograd.backward([x], [gx])
    pp.allreduce_grads(net.layers.parameters())


for _ in range(3):
    step()
torch.cuda.synchronize(); dist.barrier()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(5):
    step()
ev1.record(); torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    print(f"eager step on {world} GPUs: {ev0.elapsed_time(ev1) / 5:.2f} ms, halo rows {pp.halo.n_halo}, own {pp.n_own}, E_loc {pp.E_loc}")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); step()
    torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=32, max_name_column_width=60))
dist.destroy_process_group()
