#!/bin/bash
# Round 2, GPU call L: heavy tail reads in shared memory, extension classes sized by CTAs per SM, bm2_mem with chunks in flight:
# knob A/B, default bench, GPU tests, FASTQ->SAM and SAM benches, launch list of one step.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
W=/tmp/bm2_bench_pipe_3000_500000
( timeout 1500 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2l_bench.err | tail -1 ) > gpurun_out/r2l_bench_3gbp_1gpu.json
( timeout 900 python scripts/exp_knobs.py $W 3 2>&1 | tail -16 ) > gpurun_out/r2l_exp_knobs.log
( timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 ) > gpurun_out/r2l_tests.log 2>&1
( timeout 900 python bench.py --workload fastq2sam --steps 2 --warmup 1 2> gpurun_out/r2l_bench_fastq2sam.err | tail -1 ) > gpurun_out/r2l_bench_fastq2sam.json
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2l_launches_step.csv python scripts/prof_step.py $W 2 > /tmp/ncu_l.log 2>&1 )
cat gpurun_out/r2l_tests.log | tail -3; cat gpurun_out/r2l_exp_knobs.log | cut -c1-330
tail -c 600 gpurun_out/r2l_bench_fastq2sam.err
ls -la gpurun_out | tail -6
