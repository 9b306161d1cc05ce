import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import load_golden, rel_l2
from oracle import mgn_oracle as O
import aero_gnn_b200.models as M
DEV = "cuda:0"
n, e = int(sys.argv[1]), int(sys.argv[2])
g = load_golden("layer_sum_L2_add")
layer = M.MeshGraphNetLayer(128, 128, 128, **g["kwargs"]); layer.load_state_dict(g["state"]); layer = layer.to(DEV)
gen = torch.Generator().manual_seed(n + e)
x = torch.randn(n, 128, generator=gen); ea = torch.randn(e, 128, generator=gen)
ei = torch.randint(0, n, (2, e), generator=gen); probe = torch.randn(n + e, 128, generator=gen)
xg, eg = x.to(DEV).requires_grad_(True), ea.to(DEV).requires_grad_(True)
xo, eo = layer(xg, eg, ei.to(DEV))
(torch.cat([xo, eo], 0) * probe.to(DEV)).sum().backward()
sd = {k: v.clone().requires_grad_(True) for k, v in g["state"].items()}
xr, er = x.clone().requires_grad_(True), ea.clone().requires_grad_(True)
xo2, eo2 = O.mgn_layer(sd, "", xr, er, ei, "add")
names = list(sd)
ref = torch.autograd.grad((torch.cat([xo2, eo2], 0) * probe).sum(), [xr, er] + [sd[k] for k in names])
print("fwd", rel_l2(xo, xo2), rel_l2(eo, eo2), "g_x", rel_l2(xg.grad, ref[0]), "g_e", rel_l2(eg.grad, ref[1]))
for (k, p), r in zip(layer.named_parameters(), ref[2:]):
    print(f"{k:40s} {rel_l2(p.grad, r):.2e}")
