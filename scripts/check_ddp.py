"""torchrun --nproc-per-node N scripts/check_ddp.py : data-parallel training over batches of small meshes (C2 regime).
Every rank holds its own batch of airfoil meshes; the model is wrapped in DistributedDataParallel like any other
nn.Module.  Checks that the DDP-averaged gradients equal the average of the per-rank gradients computed without DDP
(rank 0 recomputes every rank's batch locally), fp32 exactly-ish and bf16 within bf16 tolerance, and times a step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP
import aero_gnn_b200.models as M
from aero_gnn_b200.meshes import airfoil_o_mesh, batch_meshes
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lossf = torch.nn.MSELoss()


def batch_of(r, dt):
    m = batch_meshes([airfoil_o_mesh(60, 30, seed=8 * r + s) for s in range(4)])
    return m.node_attr.to(dev, dt), m.edge_attr.to(dev, dt), m.edge_index.to(dev), m.target.to(dev)


def grads_of(net, b):
    net.zero_grad(set_to_none=True)
    loss = lossf(net(b[0], b[1], b[2]).float(), b[3])
    loss.backward()
    return [p.grad.detach().float().clone() for p in net.parameters()]


for dt, tol in ((torch.float32, 2e-4), (torch.bfloat16, 3e-2)):
    torch.manual_seed(0)
    kw = dict(bench.CFG)
    kw["processor_size"] = 3
    net = M.MeshGraphNet(6, 3, 4, **kw).to(dev).to(dt)
    ddp = DDP(net, device_ids=[local])
    mine = batch_of(rank, dt)
    ddp.zero_grad(set_to_none=True)
    loss = lossf(ddp(mine[0], mine[1], mine[2]).float(), mine[3])
    loss.backward()
    got = [p.grad.detach().float().clone() for p in net.parameters()]
    if rank == 0:
        acc = None
        for r in range(world):
            g = grads_of(net, batch_of(r, dt))
            acc = g if acc is None else [a + b for a, b in zip(acc, g)]
        want = [a / world for a in acc]
        num = sum(float((a - b).norm() ** 2) for a, b in zip(got, want)) ** 0.5
        den = sum(float(b.norm() ** 2) for b in want) ** 0.5
        err = num / den
        print(f"{dt}: DDP gradients vs average of per-rank gradients: rel L2 error {err:.3e} (tol {tol})", flush=True)
        assert err < tol
    dist.barrier()
    for _ in range(3):
        ddp.zero_grad(set_to_none=True)
        lossf(ddp(mine[0], mine[1], mine[2]).float(), mine[3]).backward()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(10):
        ddp.zero_grad(set_to_none=True)
        lossf(ddp(mine[0], mine[1], mine[2]).float(), mine[3]).backward()
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
        print(f"{dt}: {1e3 * (time.perf_counter() - t0) / 10:.2f} ms per DDP step on {world} GPUs", flush=True)
dist.barrier()
torch.cuda.synchronize()
dist.destroy_process_group()
