"""Diagnostic: per-tensor gradient errors of the tcgen05 backward vs the CUDA-core backward vs the fp32 oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import load_golden, rel_l2
from oracle import mgn_oracle as O
import aero_gnn_b200.models as M
from test_gpu_umma import _grads
DEV = "cuda:0"
name = sys.argv[1] if len(sys.argv) > 1 else "layer_sum_L2_add"
n, e = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (300, 2111)
g = load_golden(name)
layer = M.MeshGraphNetLayer(128, 128, 128, **g["kwargs"]); layer.load_state_dict(g["state"])
layer = layer.to(DEV).to(torch.bfloat16)
gen = torch.Generator().manual_seed(n + e)
x = torch.randn(n, 128, generator=gen).to(torch.bfloat16); ea = torch.randn(e, 128, generator=gen).to(torch.bfloat16)
ei = torch.randint(0, n, (2, e), generator=gen); probe = torch.randn(n + e, 128, generator=gen)
gs = _grads(layer, x.to(DEV), ea.to(DEV), ei.to(DEV), probe.to(DEV), True)
gu = _grads(layer, x.to(DEV), ea.to(DEV), ei.to(DEV), probe.to(DEV), False)
sd = {k: v.to(torch.bfloat16).float().requires_grad_(True) for k, v in g["state"].items()}
xr, er = x.float().requires_grad_(True), ea.float().requires_grad_(True)
xo, eo = O.mgn_layer(sd, "", xr, er, ei, g["kwargs"]["aggregation"])
names = list(sd)
ref = torch.autograd.grad((torch.cat([xo, eo], 0) * probe).sum(), [xr, er] + [sd[k] for k in names])
print(f"{'tensor':45s} umma-vs-ref  simt-vs-ref  umma-vs-simt")
print(f"{'g_x':45s} {rel_l2(gu[0], ref[0]):.4f}      {rel_l2(gs[0], ref[0]):.4f}      {rel_l2(gu[0], gs[0]):.4f}")
print(f"{'g_e':45s} {rel_l2(gu[1], ref[1]):.4f}      {rel_l2(gs[1], ref[1]):.4f}      {rel_l2(gu[1], gs[1]):.4f}")
for k, gr in zip(names, ref[2:]):
    print(f"{k:45s} {rel_l2(gu[2][k], gr):.4f}      {rel_l2(gs[2][k], gr):.4f}      {rel_l2(gu[2][k], gs[2][k]):.4f}")

k = "node_block.mlp.layers.0.weight"
i = names.index(k)
for nm, sl in (("W_nx", slice(0, 128)), ("W_na", slice(128, 256))):
    print(nm, "umma-vs-ref", rel_l2(gu[2][k][:, sl], ref[2 + i][:, sl]), "simt-vs-ref", rel_l2(gs[2][k][:, sl], ref[2 + i][:, sl]),
          "norms", float(ref[2 + i][:, sl].norm()))
