#!/bin/bash
# 1-GPU refresh after a kernel change: full GPU test suite, default bench line, the two ncu --set full captures.
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r02_bench_c5.json 2> $O/r02_bench_c5.err || echo "c5 bench failed"
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --graph off"
if timeout 200 $CMD > $O/plain.json 2> $O/plain.err; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_launches.csv $CMD > $O/ncu_list.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'umma_block_fwd_kernel|row_gemm_kernel' --launch-count 3 \
      -o $O/r02_fwd -f $CMD > $O/ncu_fwd.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'umma_block_bwd2_kernel|wgrad_kernel|row_gemm_kernel' \
      --launch-skip 15 --launch-count 5 -o $O/r02_bwd -f $CMD > $O/ncu_bwd.log 2>&1
fi
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 && echo "smoke ok" || echo "smoke FAILED"
ls -la $O | grep "r02_.*ncu-rep\|r02_launches"
