"""Summarise an .ncu-rep: per-kernel key metrics (raw page) and the hottest source lines (source page)."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__inst_executed.avg.per_cycle_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "launch__grid_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=" * 100)
    for k in keys:
        if k in d:
            print(f"{k:75s} {d[k]} {units[hdr.index(k)]}")
    st = sorted(((float(d[s]), s.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for s in stall), reverse=True)
    print("stalls per issue:", ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
# source page prints one CSV table per kernel, separated by headers
blocks, cur = [], []
for line in src.splitlines():
    if line.startswith('"Kernel Name"') or line.startswith('"#"') or line.startswith('"Source"'):
        if cur: blocks.append(cur)
        cur = [line]
    elif cur:
        cur.append(line)
if cur: blocks.append(cur)
for b in blocks:
    rd = list(csv.reader(b))
    h = rd[0]
    if "Source" not in h: continue
    si = h.index("Source")
    cand = [i for i, n in enumerate(h) if n in ("Warp Stall Sampling (All Samples)", "# Samples", "Samples")]
    ii = [i for i, n in enumerate(h) if n == "Instructions Executed"]
    if not cand: 
        print("columns:", h[:12]); continue
    ci = cand[0]
    tot = sum(float(r[ci] or 0) for r in rd[1:] if len(r) > ci)
    print("-" * 100, "\nhot source lines (stall samples, % of kernel):")
    best = sorted((r for r in rd[1:] if len(r) > ci and r[ci]), key=lambda r: -float(r[ci]))[:top]
    for r in best:
        inst = r[ii[0]] if ii else ""
        print(f"{100*float(r[ci])/max(tot,1):5.1f}%  inst={inst:>10s}  {r[si].strip()[:130]}")
