#!/bin/bash
# Local (no GPU): turn the files scripts/final_evidence.sh / recapture.sh brought back in gpurun_out/ into profiles/.
set -eu
cd "$(dirname "$0")/.."
B=gpurun_out/r02_bwd.ncu-rep; F=gpurun_out/r02_fwd.ncu-rep
for c in c5 c2 c3; do [ -s gpurun_out/r02_bench_$c.json ] && grep '^{' gpurun_out/r02_bench_$c.json | tail -1 > profiles/r02_bench_$c.json; done
python scripts/launch_list_summary.py gpurun_out/r02_launches.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --graph off" > profiles/r02_launch_list_summary.csv
python scripts/ncu_summary.py $B 0 > /tmp/bs.txt; python scripts/ncu_summary.py $F 0 > /tmp/fs.txt
python scripts/ncu_phase_split.py $B 2 block_umma_bwd2.cu 0,260,291,319,434,482,571,590 "prologue + weight / TMA issue lambdas,tile start (index / TMA wait),recompute H_1 H_2,output layer + LayerNorm backward,backward phases m=3..1,last phase (g_main / g_h0 out),role rotation,final flush" 100 > /tmp/ps.txt
{ echo "# ncu --set full --clock-control none --import-source on -k regex:'umma_block_bwd2_kernel|wgrad_kernel|row_gemm_kernel' --launch-skip 15 --launch-count 5"
  echo "# over: python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --graph off  (C5; layer 14 of the first backward)"
  echo "# sections: edge backward (5,996,000 rows), node backward (1,000,000 rows); hottest lines and per-phase stall split of the edge launch"
  sed -n 31,45p /tmp/bs.txt; sed -n 1,15p /tmp/bs.txt
  echo; echo "## hottest source lines, edge backward"; python scripts/ncu_lines.py $B 30 2
  echo; echo "## stall samples by phase of a tile, edge backward (scripts/ncu_phase_split.py: SASS walked in address order, attributed to"
  echo "## the last kernel-body line of block_umma_bwd2.cu seen; helper lines above the kernel inherit the phase that inlined them)"
  cat /tmp/ps.txt; } > profiles/r02_edge_bwd_ncu_summary.txt 2>&1
{ echo "# ncu --set full --clock-control none --import-source on -k regex:'umma_block_fwd_kernel|row_gemm_kernel' --launch-count 3"
  echo "# over: python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --graph off  (C5; layer 0 of the first forward)"
  sed -n 16,45p /tmp/fs.txt
  echo; echo "## hottest source lines, edge forward"; python scripts/ncu_lines.py $F 30 1; } > profiles/r02_edge_fwd_ncu_summary.txt 2>&1
{ echo "# the warp-specialised streaming kernels of a step (same two ncu --set full captures as r02_edge_{fwd,bwd}_ncu_summary.txt)"
  echo "# row_gemm P = x W^T + b (1M rows, nb = 3); wgrad<0> dW = g_h0n^T agg (1M rows); wgrad<1> dW_e = g_h0e^T e + receiver sums (6M rows); row_gemm g_x (K = 384)"
  sed -n 1,15p /tmp/fs.txt; sed -n 16,30p /tmp/bs.txt; sed -n 46,75p /tmp/bs.txt
  echo; echo "## hottest source lines, wgrad<1>"; python scripts/ncu_lines.py $B 25 3; } > profiles/r02_wgrad_rowgemm_ncu_summary.txt 2>&1
python scripts/sass_census.py > profiles/r02_sass_opcodes.txt
python scripts/ncu_traffic.py profiles/r02_ncu_traffic.json 1000000 5996000 bf16 edge_fwd=$F:1 node_fwd=$F:2 node_bwd=$B:0 edge_bwd=$B:2 > /dev/null
python -c "import bench; print('traffic lookup:', bench.ncu_traffic('edge_bwd', 1000000, 5996000, 'bf16'))"
