"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules from /root/reference on CPU.

Run here (the reference checkout does not exist on the GPU box):   python oracle/gen_golden.py
Each fixture stores: constructor kwargs, the reference state_dict, seeded inputs, the reference output,
and reference autograd gradients of  loss = sum(out * probe)  w.r.t. inputs and every parameter.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import standins  # noqa: E402

standins.install()
# the repo root also has no `models` package, so `models` resolves to the reference checkout
from models.mlp import MLP  # noqa: E402
from models.mgnLayer import EdgeBlock, EdgeBlockSum, NodeBlock, MeshGraphNetLayer  # noqa: E402
from models.mgn import MeshGraphNet  # noqa: E402
from models.bsms_mgn import BiStridedMeshGraphNet  # noqa: E402
from models.poolmgn import poolMGN  # noqa: E402
from models.fouriermgn import FourierMeshGraphNet  # noqa: E402

from aero_gnn_b200.meshes import airfoil_o_mesh, batch_meshes  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def grads_of(module, out, inputs):
    g = torch.Generator().manual_seed(99)
    probe = torch.randn(out.shape, generator=g)
    loss = (out * probe).sum()
    params = [p for p in module.parameters()]
    gs = torch.autograd.grad(loss, list(inputs) + params, allow_unused=True)
    gi = gs[: len(inputs)]
    gp = {n: (g_ if g_ is not None else torch.zeros_like(p)) for (n, p), g_ in zip(module.named_parameters(), gs[len(inputs):])}
    return probe, gi, gp


def save(name, **kw):
    torch.save(kw, os.path.join(OUT, name + ".pt"))
    print("wrote", name, {k: (tuple(v.shape) if torch.is_tensor(v) else type(v).__name__) for k, v in kw.items()})


def rand_graph(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, n, (2, e), generator=g)


def main():
    torch.manual_seed(0)
    D = 128
    # ---- one processor layer, both edge-block forms, L = 1 and 2, add and mean -------------------
    for tag, kw in {
        "layer_sum_L2_add": dict(num_hidden_layers_node_processor=2, num_hidden_layers_edge_processor=2,
                                 aggregation="add", do_concat_trick=True),
        "layer_cat_L1_mean": dict(num_hidden_layers_node_processor=1, num_hidden_layers_edge_processor=1,
                                  aggregation="mean", do_concat_trick=False),
    }.items():
        torch.manual_seed(1)
        layer = MeshGraphNetLayer(D, D, D, **kw)
        n, e = 37, 301     # ragged vs every tile size; duplicates and self-loops present
        ei = rand_graph(n, e, 5)
        g = torch.Generator().manual_seed(1234)
        x = torch.randn(n, D, generator=g, requires_grad=True)
        ea = torch.randn(e, D, generator=g, requires_grad=True)
        xo, eo = layer(x, ea, ei)
        out = torch.cat([xo, eo], 0)
        probe, gi, gp = grads_of(layer, out, [x, ea])
        save(tag, kwargs=kw, state=layer.state_dict(), x=x.detach(), e=ea.detach(), edge_index=ei, x_out=xo.detach(),
             e_out=eo.detach(), probe=probe, g_x=gi[0], g_e=gi[1], g_params=gp)

    # ---- standalone blocks ------------------------------------------------------------------------
    torch.manual_seed(2)
    n, e = 23, 97
    ei = rand_graph(n, e, 7)
    g = torch.Generator().manual_seed(4321)
    x = torch.randn(n, D, generator=g)
    ea = torch.randn(e, D, generator=g)
    eb = EdgeBlock(D, D, D, 2)
    es = EdgeBlockSum(D, D, D, 0)
    nb = NodeBlock(D, D, D, 1, aggregation="mean")
    save("blocks", state_eb=eb.state_dict(), state_es=es.state_dict(), state_nb=nb.state_dict(), x=x, e=ea,
         edge_index=ei, out_eb=eb(ea, x, ei).detach(), out_es=es(ea, x, ei).detach(), out_nb=nb(x, ea, ei).detach())

    # ---- MLP ------------------------------------------------------------------------------------------
    torch.manual_seed(3)
    m = MLP(7, 32, 16, 2, "relu")
    xin = torch.randn(11, 7)
    save("mlp", state=m.state_dict(), x=xin, out=m(xin).detach())

    # ---- full MGN on a small airfoil mesh (config.yaml kwargs, 3 steps) ----------------------------
    cfg = dict(processor_size=2, activation_fn="relu", num_hidden_layers_node_processor=2,
               num_hidden_layers_edge_processor=2, hidden_dim_processor=128, num_hidden_layers_node_encoder=2,
               hidden_dim_node_encoder=128, num_hidden_layers_edge_encoder=2, hidden_dim_edge_encoder=128,
               aggregation="add", hidden_dim_decoder=128, num_hidden_layers_decoder=2, dropout=0.0,
               do_concat_trick=True)
    torch.manual_seed(0)
    mesh = airfoil_o_mesh(12, 7, seed=0)
    net = MeshGraphNet(6, 3, 4, **cfg)
    na = mesh.node_attr.clone().requires_grad_(True)
    eattr = mesh.edge_attr.clone().requires_grad_(True)
    out = net(na, eattr, mesh.edge_index)
    probe, gi, gp = grads_of(net, out, [na, eattr])
    save("mgn", kwargs=cfg, state=net.state_dict(), node_attr=mesh.node_attr, edge_attr=mesh.edge_attr,
         edge_index=mesh.edge_index, out=out.detach(), probe=probe, g_node=gi[0], g_edge=gi[1], g_params=gp)

    # ---- BSMS on a batch of two small meshes, 3 scales --------------------------------------------
    bcfg = dict(cfg)
    bcfg.update(processor_size=5, num_scales=3, layers_per_scale=1, stride=2)
    torch.manual_seed(0)
    bm = batch_meshes([airfoil_o_mesh(10, 6, seed=0), airfoil_o_mesh(9, 5, seed=1)])
    bnet = BiStridedMeshGraphNet(6, 3, 4, **bcfg)
    na = bm.node_attr.clone().requires_grad_(True)
    eattr = bm.edge_attr.clone().requires_grad_(True)
    out = bnet(na, eattr, bm.edge_index, bm.batch, bm.pos)
    probe, gi, gp = grads_of(bnet, out, [na, eattr])
    # pooling internals of the first level (integer contract)
    with torch.no_grad():
        xh = bnet.node_encoder(bm.node_attr)
        eh = bnet.edge_encoder(bm.edge_attr)
        cx, ce, cei, cb, cpos, f2c = bnet._downsample(xh, eh, bm.edge_index, bm.batch, bm.pos)
        cx2, ce2, cei2, cb2, cpos2, f2c2 = bnet._downsample(cx, ce, cei, cb, cpos)
        _, _, cei_np, cb_np, _, f2c_np = bnet._downsample(xh, eh, bm.edge_index, bm.batch, None)
    save("bsms", kwargs=bcfg, state=bnet.state_dict(), node_attr=bm.node_attr, edge_attr=bm.edge_attr,
         edge_index=bm.edge_index, batch=bm.batch, pos=bm.pos, out=out.detach(), probe=probe, g_node=gi[0],
         g_edge=gi[1], g_params=gp, l1_f2c=f2c, l1_cei=cei, l1_cb=cb, l1_cpos=cpos, l1_cx=cx, l1_ce=ce,
         l2_f2c=f2c2, l2_cei=cei2, l2_cb=cb2, l2_cpos=cpos2, nopos_f2c=f2c_np, nopos_cei=cei_np, nopos_cb=cb_np)

    # stride 3 pooling indices on their own (ragged groups)
    b3 = BiStridedMeshGraphNet(6, 3, 4, processor_size=3, num_scales=2, layers_per_scale=1, stride=3)
    with torch.no_grad():
        _, _, cei3, cb3, cpos3, f2c3 = b3._downsample(xh, eh, bm.edge_index, bm.batch, bm.pos)
    save("bsms_stride3", edge_index=bm.edge_index, batch=bm.batch, pos=bm.pos, f2c=f2c3, cei=cei3, cb=cb3, cpos=cpos3)

    # ---- poolMGN / FourierMGN (concat edge block, L=1 defaults + config L=2) ------------------------
    pcfg = {k: v for k, v in cfg.items() if k != "do_concat_trick"}
    pcfg.update(processor_size=2)
    torch.manual_seed(0)
    pnet = poolMGN(6, 3, 4, global_pool_method="mean", num_hidden_layers_global_encoder=2, global_dim=128, **pcfg)
    out = pnet(bm.node_attr, bm.edge_attr, bm.edge_index, bm.batch)
    save("poolmgn", kwargs=pcfg, state=pnet.state_dict(), node_attr=bm.node_attr, edge_attr=bm.edge_attr,
         edge_index=bm.edge_index, batch=bm.batch, out=out.detach())
    torch.manual_seed(0)
    fnet = FourierMeshGraphNet(6, 3, 4, fourier_features_dim=2, fourier_freq_start=-3, fourier_freq_length=7, **pcfg)
    out = fnet(mesh.node_attr, mesh.edge_attr, mesh.edge_index)
    save("fouriermgn", kwargs=pcfg, state=fnet.state_dict(), node_attr=mesh.node_attr, edge_attr=mesh.edge_attr,
         edge_index=mesh.edge_index, out=out.detach(), emb=fnet.fourier_embedding(mesh.node_attr).detach())


if __name__ == "__main__":
    main()
