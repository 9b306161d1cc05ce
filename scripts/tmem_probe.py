"""Run aero_umma_probe and report (a) which TMEM lanes an M = 64 accumulator occupies, at lane offsets 0 and 16,
(b) whether the A-from-TMEM (.ts) GEMM reproduces a * b^T.  Facts for the next kernel design."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aero_gnn_b200 import lib
L = lib.load_probe()
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
a = torch.randn(128, 128, generator=g).to(dev, torch.bfloat16)
b = torch.randn(128, 128, generator=g).to(dev, torch.bfloat16)
ref = a.float() @ b.float().t()                      # [128 rows of a, 128 rows of b]
st = torch.cuda.current_stream().cuda_stream
for mode in (0, 1, 2):
    c = torch.full((128, 128), float("nan"), device=dev)
    rc = L.aero_umma_probe(a.data_ptr(), b.data_ptr(), c.data_ptr(), mode, st)
    try:
        torch.cuda.synchronize()
    except Exception as exc:                          # an illegal form traps the launch: report and stop
        print(f"mode {mode}: launch failed: {exc!r}")
        break
    if rc:
        print(f"mode {mode}: rc={rc} {L.aero_last_error().decode()}")
        continue
    if mode == 2:
        err = float((c - ref).abs().max() / ref.abs().max())
        print(f"mode 2 (A from TMEM): max rel err vs a*b^T = {err:.3e}")
        continue
    lanes = []
    for lane in range(128):
        row = c[lane]
        if float(row.abs().max()) == 0.0:
            continue
        d = (ref[:64] - row[None, :]).abs().amax(dim=1)
        r = int(d.argmin())
        lanes.append((lane, r, float(d[r] / ref.abs().max())))
    print(f"mode {mode} (M=64, D lane offset {16 * mode}): {len(lanes)} non-zero lanes")
    runs, start = [], None
    for i, (lane, r, e) in enumerate(lanes):
        if start is None:
            start = (lane, r)
        nxt = lanes[i + 1] if i + 1 < len(lanes) else None
        if nxt is None or nxt[0] != lane + 1 or nxt[1] != r + 1:
            runs.append((start[0], lane, start[1], r))
            start = None
    for l0, l1, r0, r1 in runs:
        print(f"   TMEM lanes {l0:3d}..{l1:3d}  <-  rows {r0:2d}..{r1:2d} of the 64-row product")
    print("   worst match error:", max((e for _, _, e in lanes), default=float('nan')))

# ---- MMA rate per operand orientation ----
out2 = torch.zeros(2, dtype=torch.int64, device=dev)
for a_mn, b_mn, what in ((0, 0, "A K-major, B K-major (forward)"), (0, 1, "A K-major, B MN-major (data gradient)"),
                         (1, 1, "A MN-major, B MN-major (weight gradient)")):
    for reps in (1, 4, 32):
        L.aero_umma_rate_probe(a.data_ptr(), b.data_ptr(), out2.data_ptr(), a_mn, b_mn, reps, st)
        torch.cuda.synchronize()
        cyc, n = int(out2[0]), int(out2[1])
        print(f"{what}: {n:4d} MMAs (128x128x16) in {cyc:6d} cycles = {cyc / n:6.1f} cycles/MMA")
