#!/bin/bash
# Round 2, GPU call S: the warp post-filter with four 32-box steps classified at a time; which reads set the duration of the warp-per-read tail
# kernel (BM2_DEBUG_NREG) next to the kernel's duration on the same box.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
W=/tmp/bm2_bench_pipe_3000_500000
( timeout 1500 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2s_bench.err | tail -1 ) > gpurun_out/r2s_bench_3gbp_1gpu.json
( timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 ) > gpurun_out/r2s_tests.log 2>&1
( timeout 900 python scripts/exp_knobs.py $W 3 2>&1 | tail -8 ) > gpurun_out/r2s_exp_knobs.log
( BM2_DEBUG_NREG=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tail_kernel|chain_kernel' -c 16 --csv --log-file gpurun_out/r2s_tail_launches.csv python scripts/prof_step.py $W 2 2>&1 | grep "bm2 debug" | head -4 ) > gpurun_out/r2s_heaviest_reads.log
cat gpurun_out/r2s_tests.log | tail -3; cat gpurun_out/r2s_heaviest_reads.log; grep -i "tail_kernel\|chain_kernel" gpurun_out/r2s_tail_launches.csv | cut -c1-60,200-400 | tail -8; cat gpurun_out/r2s_exp_knobs.log | cut -c1-200
ls -la gpurun_out | tail -6
