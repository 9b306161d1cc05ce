"""torchrun --nproc-per-node N scripts/check_partition.py : the receiver-block partitioned processor (forward,
input gradients, weight gradients) equals the single-GPU processor on the same mesh and weights."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import aero_gnn_b200.models as M
from aero_gnn_b200 import ops
from aero_gnn_b200.meshes import wing_surface_mesh
from aero_gnn_b200.models._common import run_layers
from aero_gnn_b200.partition import PartitionedProcessor

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
dt = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float32
mesh = wing_surface_mesh(60, 40)
N, E = mesh.num_nodes, mesh.num_edges
torch.manual_seed(0)
kw = dict(processor_size=3, num_hidden_layers_node_processor=2, num_hidden_layers_edge_processor=2,
          aggregation="add", do_concat_trick=True)
net = M.MeshGraphNet(6, 4, 5, **kw).to(dev).to(dt)
g = torch.Generator().manual_seed(7)
xg = torch.randn(N, 128, generator=g).to(dt)
eg = torch.randn(E, 128, generator=g).to(dt)          # caller edge order
probe = torch.randn(N, 128, generator=g)

# single-GPU reference on this rank
plan = ops.PLAN_CACHE.get(mesh.edge_index.to(dev), N)
x0 = xg.to(dev).requires_grad_(True)
e0 = eg.to(dev)[plan.perm.long()].requires_grad_(True)
xr, _ = run_layers(net.layers, plan, x0, e0)
(xr.float() * probe.to(dev)).sum().backward()
ref_gx, ref_ge = x0.grad.clone(), e0.grad.clone()
ref_gw = [p.grad.clone() for p in net.layers.parameters()]
for p in net.layers.parameters():
    p.grad = None

pp = PartitionedProcessor(mesh.edge_index, N, rank, world, dev)
x1 = xg[pp.lo:pp.hi].to(dev).requires_grad_(True)
ids = pp.csr_edge_ids()
e1 = eg.to(dev)[ids].requires_grad_(True)
xo, _ = pp.run(net.layers, x1, e1)
(xo.float() * probe[pp.lo:pp.hi].to(dev)).sum().backward()
pp.allreduce_grads(net.layers.parameters())

def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30))
tol = 1e-5 if dt == torch.float32 else 2e-2
errs = {"x": rel(xo, xr[pp.lo:pp.hi]), "g_x": rel(x1.grad, ref_gx[pp.lo:pp.hi])}
# local CSR slot k holds global caller edge ids[k], which the single-GPU plan keeps at CSR slot inv_perm[ids[k]]
errs["g_e"] = rel(e1.grad, ref_ge[plan.inv_perm.long()[ids]])
errs["g_w"] = max(rel(p.grad, r) for p, r in zip(net.layers.parameters(), ref_gw))
errs["interior_edges"] = pp.E_int / max(pp.E_loc, 1)
bad = {k: v for k, v in errs.items() if k != "interior_edges" and not v < tol}
print(f"rank {rank}/{world} dtype {dt} n_own {pp.n_own} n_halo {pp.halo.n_halo} E_loc {pp.E_loc} errs {errs}", flush=True)

# ---- GMP / WeightedEdgeConv variant (BFS-bistride operators) on the same partition ------------------------------
if dt == torch.float32:
    torch.manual_seed(1)
    gmps = torch.nn.ModuleList(M.GMP(128, 128, 128) for _ in range(2)).to(dev)
    conv = M.WeightedEdgeConv(128, 128).to(dev)
    pos = mesh.pos[:, :2].clone() + 1e-3 * torch.rand(N, 2, generator=torch.Generator().manual_seed(3))
    ei_d = mesh.edge_index.to(dev)
    # single GPU
    x0 = xg.to(dev).requires_grad_(True)
    e0 = eg.to(dev).requires_grad_(True)
    x, e = x0, e0
    for m_ in gmps:
        x, e = m_(x, e, ei_d)
    oc, wc = conv(x, ei_d, pos.to(dev))
    (oc * probe.to(dev)).sum().backward()
    ref2 = dict(out=oc.detach(), w=wc.detach(), gx=x0.grad.clone(), ge=e0.grad.clone(),
                gw=[p.grad.clone() for p in list(gmps.parameters()) + list(conv.parameters())])
    for p in list(gmps.parameters()) + list(conv.parameters()):
        p.grad = None
    # partitioned: GMP stack through pp.run, WeightedEdgeConv on the halo-extended rows, positions exchanged once
    pos_ext = pp.extend_static(pos[pp.lo:pp.hi].to(dev))
    x1 = xg[pp.lo:pp.hi].to(dev).requires_grad_(True)
    e1 = eg.to(dev)[ids].requires_grad_(True)
    xo2, _ = pp.run(gmps, x1, e1)
    oc2, wc2 = pp.weighted_edge_conv(conv, xo2, pos_ext)
    (oc2 * probe[pp.lo:pp.hi].to(dev)).sum().backward()
    others = list(conv.parameters())
    pp.allreduce_grads(others)
    eids = torch.from_numpy(pp.halo.edge_ids).to(dev)
    errs2 = {"wec_out": rel(oc2, ref2["out"][pp.lo:pp.hi]), "wec_w": rel(wc2, ref2["w"][eids]),
             "g_x": rel(x1.grad, ref2["gx"][pp.lo:pp.hi]), "g_e": rel(e1.grad, ref2["ge"][ids]),
             "g_w": max(rel(p.grad, r) for p, r in zip(list(gmps.parameters()) + others, ref2["gw"]))}
    print(f"rank {rank}/{world} GMP+WeightedEdgeConv on the partition: {errs2}", flush=True)
    bad.update({"gmp_" + k: v for k, v in errs2.items() if not v < 1e-4})
t = torch.tensor([len(bad)], device=dev)
dist.all_reduce(t)
dist.destroy_process_group()
sys.exit(1 if int(t.item()) else 0)
