"""BFS-bistride hierarchy + WeightedEdgeConv kernels on one B200: time per kernel (CUDA events on the launching stream,
median of 20 after 5 warm-ups, L2 flushed between iterations) and the HBM fraction of the algorithmic traffic.

    python scripts/bistride_bench.py [c3|c5] [bf16|fp32]

c3: airfoil O-mesh 400x250 (N=100,000, E=598,400); c5: wing surface 1000x1000 (N=1,000,000, E=5,996,000).
Algorithmic bytes of WeightedEdgeConv (b = bytes per element, out = in = 128, hidden 64):
  forward : N*(256+128... read Q [N,256] once, write out [N,128], write w [E], read src/perm [E] int32, rowptr, pos
  backward: read Q, g_out; write dQ [N,256]; read w [E], ds write+read [E] fp32; src/dst/perm/sperm [E] int32
Gathered rows (A[src], T[src], B[dst], g_out[dst]) are re-reads of those matrices and are expected to hit L2.
"""
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import aero_gnn_b200.models as M
from aero_gnn_b200 import bistride as bs, ops
from aero_gnn_b200.meshes import airfoil_o_mesh, wing_surface_mesh

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
dt = torch.bfloat16 if (len(sys.argv) < 3 or sys.argv[2] == "bf16") else torch.float32
PROF = len(sys.argv) > 3 and sys.argv[3] == "prof"    # one launch of each kernel, no warm-up: for an ncu capture
dev = torch.device("cuda", 0)
mesh = airfoil_o_mesh(400, 250, seed=0) if which == "c3" else wing_surface_mesh(1000, 1000)
pos = mesh.pos[:, :2].contiguous().to(dev)
ei = mesh.edge_index.to(dev)
N, E = mesh.num_nodes, mesh.num_edges
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6458.7))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=20, warm=5):
    if PROF:
        n, warm = 1, 0
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


res = {"mesh": which, "N": N, "E": E, "dtype": str(dt), "hbm_peak_gbps": hbm}
plan = ops.PLAN_CACHE.get(ei, N)
seed = bs.seed_node(ei, N, pos, plan)
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
d = bs.bfs_levels(plan, seed)
torch.cuda.synchronize(); t_bfs = (time.perf_counter() - t0) * 1e3
res["bfs"] = {"levels": int(d.max()) + 1, "wall_ms": t_bfs, "launches": ops.LaunchCounter.total}
ms_sel = timed(lambda: bs.bistride_select(d))
sel, imap, fb = bs.bistride_select(d)
ms_fil = timed(lambda: bs.filter_edges(ei, imap))
cei, kept = bs.filter_edges(ei, imap)
res["select"] = {"ms": ms_sel, "n_selected": int(sel.numel()), "fallback": fb,
                 "alg_gbps": N * (8 + 8 + 8) / ms_sel / 1e6}
res["filter_edges"] = {"ms": ms_fil, "coarse_edges": int(cei.size(1)), "alg_gbps": (E * 16 + cei.numel() * 8) / ms_fil / 1e6}
t0 = time.perf_counter()
multi = M.MultiScaleGraphPreprocessor(3).create_multiscale_graph(types.SimpleNamespace(edge_index=ei, pos=pos))
torch.cuda.synchronize()
res["hierarchy_3_levels"] = {"wall_ms": (time.perf_counter() - t0) * 1e3, "num_nodes": multi["num_nodes"],
                             "num_edges": [int(e.size(1)) for e in multi["edge_indices"]]}

# ---- WeightedEdgeConv kernels alone (descriptor-level, GEMMs excluded) -------------------------------------------
import ctypes as C
from aero_gnn_b200 import lib as L
lib = L.load()
b = 2 if dt == torch.bfloat16 else 4
torch.manual_seed(0)
Q = torch.randn(N, 256, device=dev).to(dt)
w1l, w2, b2 = torch.randn(64, device=dev) * 0.1, torch.randn(64, device=dev) * 0.1, torch.zeros(1, device=dev)
w = torch.empty(E, 1, dtype=dt, device=dev)
out = torch.empty(N, 128, dtype=dt, device=dev)
g_out = torch.randn(N, 128, device=dev).to(dt)
dQ = torch.empty_like(Q)
g_small = torch.zeros(129, device=dev)
pos32 = pos.float().contiguous()
desc = bs._wec_desc(plan, Q, 128, False, True, pos32, w1l, w2, b2, w)
desc.out, desc.g_out, desc.dQ, desc.g_small = out.data_ptr(), g_out.data_ptr(), dQ.data_ptr(), g_small.data_ptr()
ws = torch.empty(lib.aero_wec_workspace_bytes(C.byref(desc), 1), dtype=torch.uint8, device=dev)
desc.workspace, desc.workspace_bytes = ws.data_ptr(), ws.numel()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
ms_f = timed(lambda: L.check(lib.aero_wec_fwd(C.byref(desc), st), "fwd"))
ms_b = timed(lambda: L.check(lib.aero_wec_bwd(C.byref(desc), st), "bwd"))
alg_f = N * 256 * b + N * 128 * b + E * b + E * 8 + N * 4 + N * 8
alg_b = N * 256 * b + N * 128 * b + N * 256 * b + E * b + E * 8 + E * 16 + N * 8 + N * 8
gath_f = E * (64 + 128) * b
gath_b = E * (64 + 128) * b + E * (64 + 128) * b
res["wec_fwd"] = {"ms": ms_f, "alg_bytes": alg_f, "alg_gbps": alg_f / ms_f / 1e6, "frac_hbm": alg_f / ms_f / 1e6 / hbm,
                  "gathered_bytes": gath_f, "gather_gbps": (alg_f + gath_f) / ms_f / 1e6, "edges_per_s": E / ms_f * 1e3}
res["wec_bwd"] = {"ms": ms_b, "alg_bytes": alg_b, "alg_gbps": alg_b / ms_b / 1e6, "frac_hbm": alg_b / ms_b / 1e6 / hbm,
                  "gathered_bytes": gath_b, "gather_gbps": (alg_b + gath_b) / ms_b / 1e6, "edges_per_s": E / ms_b * 1e3}

# ---- module level: WeightedEdgeConv and GMP forward + backward through the public API ---------------------------
conv = M.WeightedEdgeConv(128, 128).to(dev).to(dt)
gmp = M.GMP(128, 128, 128).to(dev).to(dt)
x = torch.randn(N, 128, device=dev).to(dt).requires_grad_(True)
ea = torch.randn(E, 128, device=dev).to(dt).requires_grad_(True)


def conv_step():
    o, ww = conv(x, ei, pos)
    o.backward(g_out)


def gmp_step():
    xo, eo = gmp(x, ea, ei)
    xo.backward(g_out)


res["wec_module_fwd_bwd_ms"] = timed(conv_step, n=10, warm=3)
res["gmp_module_fwd_bwd_ms"] = timed(gmp_step, n=10, warm=3)
print(json.dumps(res))
