#!/bin/bash
# Round 2, GPU call C: BSW with loads one pair ahead + (qlen, h0) job order; sharded start-up on the GPU; first long-read numbers;
# source-level counters of the biggest bsw_col2 launch.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
W=/tmp/bm2_bench_pipe_3000_500000
( timeout 600 python -m pytest tests/test_bsw_gpu.py tests/test_shard_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/r2c_tests.log 2>&1
( timeout 1200 python bench.py --steps 5 --warmup 3 2> gpurun_out/r2c_bench.err | tail -1 ) > gpurun_out/r2c_bench_3gbp_1gpu.json
( timeout 900 python bench.py --workload longread --long-reads 512 --long-sample 512 --steps 1 --warmup 1 2> gpurun_out/r2c_bench_long.err | tail -1 ) > gpurun_out/r2c_bench_long.json
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:bsw_col2 -s 21 -c 1 -o /tmp/r2c_bsw python scripts/prof_step.py $W 2 > /tmp/ncu_bsw.log 2>&1 ;
  [ -f /tmp/r2c_bsw.ncu-rep ] && python scripts/ncu_summary.py /tmp/r2c_bsw.ncu-rep gpurun_out/r2c_bsw_col2_one.md 'bsw_col2_kernel, right extensions of the 128-column class' &&
  ncu -i /tmp/r2c_bsw.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2c_src_bsw_col2.csv.gz ) > gpurun_out/r2c_ncu_bsw.log 2>&1
tail -c 400 gpurun_out/r2c_bench_long.err; cat gpurun_out/r2c_tests.log | tail -3
ls -la gpurun_out | tail -12
