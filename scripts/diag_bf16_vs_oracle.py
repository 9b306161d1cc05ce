"""bf16 prediction error of every model family on its BASELINE.json config against the fp32 CPU oracle evaluated on the
bf16-held parameters / inputs, next to the error of the reference's own bf16 mode (the same oracle restatement run as
pure bf16 torch ops on the GPU, train.py:30-33) against the same truth.  Run on the GPU box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aero_gnn_b200.models as M
from aero_gnn_b200.meshes import airfoil_o_mesh, batch_meshes
from oracle import mgn_oracle as O

DEV = "cuda:0"


def rrmse(pred, ref):
    pred, ref = pred.detach().double().cpu(), ref.detach().double().cpu()
    rmse = ((pred - ref) ** 2).mean(0).sqrt()
    sc = ref.abs().mean(0)
    return float(torch.where(sc > 1e-8, rmse / sc, torch.zeros_like(rmse)).mean())


def rl2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm())


base = dict(processor_size=15, num_hidden_layers_node_processor=2, num_hidden_layers_edge_processor=2,
            num_hidden_layers_node_encoder=2, num_hidden_layers_edge_encoder=2, num_hidden_layers_decoder=2,
            aggregation="add")
c2 = batch_meshes([airfoil_o_mesh(100, 50, seed=s) for s in range(8)])
c3 = airfoil_o_mesh(400, 250, seed=0)
cases = [
    ("mgn_c2", c2, lambda: M.MeshGraphNet(6, 3, 4, do_concat_trick=True, **base),
     lambda net, m, a, b: net(a, b, m.edge_index.to(a.device)),
     lambda sd, m, a, b: O.mgn_forward(sd, a, b, m.edge_index.to(a.device))),
    ("mgn_100k", c3, lambda: M.MeshGraphNet(6, 3, 4, do_concat_trick=True, **base),
     lambda net, m, a, b: net(a, b, m.edge_index.to(a.device)),
     lambda sd, m, a, b: O.mgn_forward(sd, a, b, m.edge_index.to(a.device))),
    ("bsms", c3, lambda: M.BiStridedMeshGraphNet(6, 3, 4, do_concat_trick=True, num_scales=4, layers_per_scale=2, stride=2, **base),
     lambda net, m, a, b: net(a, b, m.edge_index.to(a.device), m.batch.to(a.device), m.pos.to(a.device)),
     lambda sd, m, a, b: O.bsms_forward(sd, a, b, m.edge_index.to(a.device), m.batch.to(a.device), m.pos.to(a.device).to(a.dtype))),
    ("poolmgn", c3, lambda: M.poolMGN(6, 3, 4, global_pool_method="mean", num_hidden_layers_global_encoder=2, global_dim=128, **base),
     lambda net, m, a, b: net(a, b, m.edge_index.to(a.device), m.batch.to(a.device)),
     lambda sd, m, a, b: O.pool_mgn_forward(sd, a, b, m.edge_index.to(a.device), m.batch.to(a.device))),
    ("fourier", c3, lambda: M.FourierMeshGraphNet(6, 3, 4, **base),
     lambda net, m, a, b: net(a, b, m.edge_index.to(a.device)),
     lambda sd, m, a, b: O.fourier_mgn_forward(sd, a, b, m.edge_index.to(a.device))),
]
only = sys.argv[1:] or [c[0] for c in cases]
for name, mesh, make, call, oracle in cases:
    if name not in only:
        continue
    torch.manual_seed(0)
    net16 = make().to(DEV).to(torch.bfloat16)
    na16, ea16 = mesh.node_attr.to(torch.bfloat16), mesh.edge_attr.to(torch.bfloat16)
    sd32 = {k: v.detach().float().cpu() for k, v in net16.state_dict().items()}
    t0 = time.time()
    with torch.no_grad():
        truth = oracle(sd32, mesh, na16.float(), ea16.float())                       # fp32 CPU oracle, bf16-held values
    t1 = time.time()
    with torch.no_grad():
        ours = call(net16, mesh, na16.to(DEV), ea16.to(DEV)).float().cpu()
        sd16 = {k: v.detach() for k, v in net16.state_dict().items()}
        ref16 = oracle(sd16, mesh, na16.to(DEV), ea16.to(DEV)).float().cpu()        # reference's own bf16 mode
    print(f"{name:9s} ours: rrmse {rrmse(ours, truth):.4f} rel-L2 {rl2(ours, truth):.4f} | reference bf16 mode: rrmse "
          f"{rrmse(ref16, truth):.4f} rel-L2 {rl2(ref16, truth):.4f} | oracle {t1 - t0:.1f}s", flush=True)
