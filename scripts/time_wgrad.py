"""Times aero_wgrad against the library GEMM + aero_segment_reduce it replaces, at the C5 sizes (CUDA events)."""
import torch
from aero_gnn_b200 import ops

dev = "cuda:0"
N, E = 1_000_000, 5_996_000


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


ei = torch.stack([torch.randint(0, N, (E,)), torch.arange(E) % N]).to(dev)
plan = ops.build_graph_plan(ei, N)
G = torch.randn(E, 128, device=dev).bfloat16()
X = torch.randn(E, 128, device=dev).bfloat16()
out = torch.empty(128, 128, device=dev)
psd = torch.empty(N, 256, device=dev, dtype=torch.bfloat16)
gb = 2 * E * 256 / 1e9
t = timeit(lambda: ops.wgrad(G, X, out)); print(f"wgrad a=1 E rows           {t:.3f} ms  {gb / t * 1e3:.0f} GB/s")
t = timeit(lambda: ops.wgrad(G, X, out, seg=(plan.dst, plan.rowptr, N, psd[:, 128:]))); print(f"wgrad a=1 E rows + seg     {t:.3f} ms  {gb / t * 1e3:.0f} GB/s")
t = timeit(lambda: torch.mm(G.t(), X, out_dtype=torch.float32, out=out)); print(f"torch.mm E rows            {t:.3f} ms  {gb / t * 1e3:.0f} GB/s")
t = timeit(lambda: ops.segment_reduce(G, plan.rowptr, None, N, out=psd[:, 128:])); print(f"segment_reduce (receiver)  {t:.3f} ms")
t = timeit(lambda: ops.segment_reduce(G, plan.sptr, plan.sperm, N, out=psd[:, :128])); print(f"segment_reduce (sender)    {t:.3f} ms")
Gn, Xn = G[:N].contiguous(), X[:N].contiguous()
out2 = torch.empty(256, 128, device=dev)
psd.normal_()
t = timeit(lambda: ops.wgrad(psd, Xn, out2)); print(f"wgrad a=2 N rows           {t:.3f} ms  {N * 768 / 1e6 / t:.0f} GB/s")
t = timeit(lambda: torch.mm(psd.t(), Xn, out_dtype=torch.float32, out=out2)); print(f"torch.mm a=2 N rows        {t:.3f} ms")
t = timeit(lambda: ops.wgrad(Gn, Xn, out)); print(f"wgrad a=1 N rows           {t:.3f} ms  {N * 512 / 1e6 / t:.0f} GB/s")
t = timeit(lambda: torch.mm(Gn.t(), Xn, out_dtype=torch.float32, out=out)); print(f"torch.mm a=1 N rows        {t:.3f} ms")

from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        ops.wgrad(G, X, out, seg=(plan.dst, plan.rowptr, N, psd[:, 128:]))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=6, max_name_column_width=60))

# row GEMMs of a step: projection P and back-projection g_x, against the library calls they replace
Wp = (torch.randn(384, 128, device=dev) / 11).bfloat16()
bp = torch.randn(384, device=dev).bfloat16()
Gx = torch.randn(N, 128, device=dev).bfloat16()
t = timeit(lambda: ops.row_gemm([Xn], Wp, w_mn=False, nb=3, bias=bp)); print(f"row_gemm P (N rows)        {t:.3f} ms  {N * 1024 / 1e6 / t:.0f} GB/s")
t = timeit(lambda: torch.addmm(bp, Xn, Wp.t())); print(f"torch.addmm P              {t:.3f} ms")
t = timeit(lambda: ops.row_gemm([psd[:, :128], psd[:, 128:], Gn], Wp, w_mn=True, add=Gx)); print(f"row_gemm g_x (K=384)       {t:.3f} ms  {N * 1280 / 1e6 / t:.0f} GB/s")
def lib_gx():
    g = torch.addmm(Gx, psd, Wp[:256])
    g.addmm_(Gn, Wp[256:])
t = timeit(lib_gx); print(f"torch addmm + addmm_ g_x   {t:.3f} ms")
