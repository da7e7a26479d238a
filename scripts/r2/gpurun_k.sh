#!/bin/bash
# Round 2, GPU call K: unique-interval text shortcut (SMEM), per-warp dynamic job fetch + implicit-column row maximum (BSW), chain coop threshold:
# whole GPU test suite, default bench (parity against the reference's regs), A/B of the knobs at 1 and 4 sub-batches, ncu of one extension launch.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
W=/tmp/bm2_bench_pipe_3000_500000
( timeout 1500 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2k_bench.err | tail -1 ) > gpurun_out/r2k_bench_3gbp_1gpu.json
( timeout 900 python scripts/exp_knobs.py $W 3 2>&1 | tail -14 ) > gpurun_out/r2k_exp_knobs.log
( timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 ) > gpurun_out/r2k_tests.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bsw_col2_kernel' -s 9 -c 1 -o /tmp/r2k_bsw python scripts/prof_step.py $W 1 > /tmp/ncu_bsw.log 2>&1 ;
  [ -f /tmp/r2k_bsw.ncu-rep ] && python scripts/ncu_summary.py /tmp/r2k_bsw.ncu-rep gpurun_out/r2k_bsw_col2.md 'bsw_col2_kernel, 128-column class, per-warp job fetch + implicit-column key' &&
  ncu -i /tmp/r2k_bsw.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2k_src_bsw_col2.csv.gz ) > gpurun_out/r2k_ncu_bsw.log 2>&1
cat gpurun_out/r2k_tests.log | tail -3; cat gpurun_out/r2k_exp_knobs.log | cut -c1-400
ls -la gpurun_out | tail -6
