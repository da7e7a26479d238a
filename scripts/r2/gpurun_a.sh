#!/bin/bash
# Round 2, GPU call A: suite, default bench (new bench.py: parity vs the reference's regs, CPU arm in one process, Occ device layout),
# layout A/B + SMEM CTA sweep, the SAM stage's first timing, launch list of the SAM stage.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
( timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 ) > gpurun_out/r2a_tests_gpu.log 2>&1
( timeout 1200 python bench.py --steps 5 --warmup 3 2> gpurun_out/r2a_bench.err | tail -1 ) > gpurun_out/r2a_bench_3gbp_1gpu.json
( timeout 600 python scripts/r2/exp_layout.py /tmp/bm2_bench_pipe_3000_500000 2 2>&1 | tail -20 ) > gpurun_out/r2a_exp_layout.log
( timeout 900 python bench.py --workload sam --steps 2 --warmup 1 2> gpurun_out/r2a_bench_sam.err | tail -1 ) > gpurun_out/r2a_bench_sam.json
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'sam_|ksw_' -c 60 --csv --log-file gpurun_out/r2a_launches_sam.csv \
    python bench.py --workload sam --steps 1 --warmup 1 > /tmp/ncu_sam.log 2>&1 )
tail -c 600 gpurun_out/r2a_bench.err
ls -la gpurun_out | tail -12
