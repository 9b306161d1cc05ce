import sys; sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
import aero_gnn_b200.models as M
from aero_gnn_b200.meshes import airfoil_o_mesh
from conftest import rrmse
DEV = torch.device("cuda", 0)
mesh = airfoil_o_mesh(400, 250, seed=0)
na, ea, ei, tg = mesh.node_attr.to(DEV), mesh.edge_attr.to(DEV), mesh.edge_index.to(DEV), mesh.target.to(DEV)
batch = mesh.batch.to(DEV)
base = dict(processor_size=15, num_hidden_layers_node_processor=2, num_hidden_layers_edge_processor=2,
            num_hidden_layers_node_encoder=2, num_hidden_layers_edge_encoder=2, num_hidden_layers_decoder=2, aggregation="add")
cases = [("mgn", lambda: M.MeshGraphNet(6, 3, 4, do_concat_trick=True, **base), lambda n, a, b: n(a, b, ei)),
         ("mgn_cat", lambda: M.MeshGraphNet(6, 3, 4, do_concat_trick=False, **base), lambda n, a, b: n(a, b, ei)),
         ("fourier", lambda: M.FourierMeshGraphNet(6, 3, 4, **base), lambda n, a, b: n(a, b, ei)),
         ("poolmgn", lambda: M.poolMGN(6, 3, 4, global_pool_method="mean", num_hidden_layers_global_encoder=2, global_dim=128, **base), lambda n, a, b: n(a, b, ei, batch))]
for name, make, call in cases:
    torch.manual_seed(0)
    net16 = make().to(DEV).to(torch.bfloat16)
    net32 = make().to(DEV)
    net32.load_state_dict({k: v.float() for k, v in net16.state_dict().items()})
    with torch.no_grad():
        ref = call(net32, na.to(torch.bfloat16).float(), ea.to(torch.bfloat16).float())
        out = call(net16, na.to(torch.bfloat16), ea.to(torch.bfloat16))
    e = out.float() - ref
    print(name, "RRMSE", round(rrmse(out.float(), ref), 5), "rel L2", round(float(e.norm() / ref.norm()), 5), "per-feature RMSE", [round(float(v), 5) for v in e.pow(2).mean(0).sqrt()], "per-feature mean abs ref", [round(float(v), 4) for v in ref.abs().mean(0)])
