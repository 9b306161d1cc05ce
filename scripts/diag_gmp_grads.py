"""fp32 gradient diagnosis on a random multigraph: GMP (L=0 blocks) and the golden MeshGraphNetLayer (L=2) against the
fp64 oracle; prints per-tensor errors and the degrees of the worst node rows.  usage: diag_gmp_grads.py N E"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from conftest import load_golden, rel_l2, rel_err
from oracle import mgn_oracle as O, bistride_oracle as B
import aero_gnn_b200.models as M
DEV = "cuda:0"
n, e = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(6)
ei = torch.from_numpy(rng.integers(0, n, size=(2, e)).astype(np.int64))
g = torch.Generator().manual_seed(5)
x, ea = torch.randn(n, 128, generator=g), torch.randn(e, 128, generator=g)
gx, ge = torch.randn(n, 128, generator=g), torch.randn(e, 128, generator=g)
indeg = torch.bincount(ei[1], minlength=n); outdeg = torch.bincount(ei[0], minlength=n)


def report(tag, got, ref, names):
    for k, a, r in zip(names, got, ref):
        print(f"  {tag} {k:34s} rel_l2 {rel_l2(a, r):.2e}  max {rel_err(a, r):.2e}")
    d = (got[0].detach().double().cpu() - ref[0].double()).abs().max(dim=1).values
    worst = torch.topk(d, 5).indices
    print("  worst g_x rows:", [(int(i), float(d[i]), int(indeg[i]), int(outdeg[i])) for i in worst],
          " zero-in nodes:", int((indeg == 0).sum()), " zero-out:", int((outdeg == 0).sum()))
    d = (got[1].detach().double().cpu() - ref[1].double()).abs().max(dim=1).values
    worst = torch.topk(d, 5).indices
    print("  worst g_e rows:", [(int(i), float(d[i]), int(ei[0, i]), int(ei[1, i])) for i in worst])


for which in ("gmp", "layer"):
    torch.manual_seed(4)
    if which == "gmp":
        mod = M.GMP(128, 128, 128)
        sd = {k: v.detach().clone() for k, v in mod.state_dict().items()}
        fwd = lambda s, a, b: B.gmp(s, "", a, b, ei)
    else:
        gg = load_golden("layer_sum_L2_add")
        mod = M.MeshGraphNetLayer(128, 128, 128, **gg["kwargs"]); mod.load_state_dict(gg["state"])
        sd = {k: v.detach().clone() for k, v in mod.state_dict().items()}
        fwd = lambda s, a, b: O.mgn_layer(s, "", a, b, ei, "add")
    mod = mod.to(DEV)
    xd, ed = x.to(DEV).requires_grad_(True), ea.to(DEV).requires_grad_(True)
    xo, eo = mod(xd, ed, ei.to(DEV))
    torch.autograd.backward([xo, eo], [gx.to(DEV), ge.to(DEV)])
    names = list(sd)
    s64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    x64, e64 = x.double().requires_grad_(True), ea.double().requires_grad_(True)
    xo64, eo64 = fwd(s64, x64, e64)
    ref = torch.autograd.grad([xo64, eo64], [x64, e64] + [s64[k] for k in names], [gx.double(), ge.double()])
    print(which, "fwd", rel_err(xo, xo64), rel_err(eo, eo64))
    got = [xd.grad, ed.grad] + [dict(mod.named_parameters())[k].grad for k in names]
    report(which, got, ref, ["x", "e"] + names)
    # only-x-gradient and only-e-gradient runs separate the two residual streams
    for tag, seeds in (("G_x only", (gx, None)), ("G_e only", (None, ge))):
        xd.grad = ed.grad = None
        xo, eo = mod(xd, ed, ei.to(DEV))
        outs, gs = ([xo], [seeds[0].to(DEV)]) if seeds[1] is None else ([eo], [seeds[1].to(DEV)])
        torch.autograd.backward(outs, gs)
        xo64, eo64 = fwd(s64, x64, e64)
        o64, g64 = ([xo64], [seeds[0].double()]) if seeds[1] is None else ([eo64], [seeds[1].double()])
        r = torch.autograd.grad(o64, [x64, e64], g64, allow_unused=True)
        print(f"  {tag}: g_x max {rel_err(xd.grad, r[0]):.2e}  g_e max {rel_err(ed.grad, r[1]):.2e}")
