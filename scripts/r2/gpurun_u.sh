#!/bin/bash
# Round 2, GPU call U: source-level profiles of the two warp-per-read kernels (tail, chain) of the final tree; the GPU suite once more.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
W=/tmp/bm2_bench_pipe_3000_500000
( timeout 1500 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2u_bench.err | tail -1 ) > gpurun_out/r2u_bench_3gbp_1gpu.json
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tail_kernel|chain_kernel' -s 4 -c 4 -o /tmp/r2u_ct python scripts/prof_step.py $W 2 > /tmp/ncu_u.log 2>&1 ;
  [ -f /tmp/r2u_ct.ncu-rep ] && python scripts/ncu_summary.py /tmp/r2u_ct.ncu-rep gpurun_out/r2u_chain_tail.md 'chain_kernel (light, heavy) and tail_kernel<0>, <1> of one unsplit step, final tree' &&
  ncu -i /tmp/r2u_ct.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2u_src_chain_tail.csv.gz ) > gpurun_out/r2u_ncu.log 2>&1
( timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 ) > gpurun_out/r2u_tests.log 2>&1
cat gpurun_out/r2u_tests.log | tail -3; tail -3 gpurun_out/r2u_ncu.log
ls -la gpurun_out | tail -6
