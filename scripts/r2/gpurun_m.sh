#!/bin/bash
# Round 2, GPU call M: light reads of the chain / tail kernels in work order, pair loop unrolled x8 (A/B), issue-slot use of the extension
# launches by class (warps per SM differ by class: what more latency hiding would be worth), quick parity tests.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
W=/tmp/bm2_bench_pipe_3000_500000
( timeout 1500 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2m_bench.err | tail -1 ) > gpurun_out/r2m_bench_3gbp_1gpu.json
( timeout 900 python scripts/exp_knobs.py $W 3 2>&1 | tail -12 ) > gpurun_out/r2m_exp_knobs.log
( timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_bsw_gpu.py tests/test_longreads_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 ) > gpurun_out/r2m_tests.log 2>&1
( timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_alu.sum,smsp__thread_inst_executed_per_inst_executed.ratio,launch__occupancy_limit_shared_mem,launch__grid_size \
    --clock-control none -k regex:'bsw_col2_kernel' -c 48 --csv --log-file gpurun_out/r2m_bsw_by_class.csv python scripts/prof_step.py $W 2 > /tmp/ncu_m.log 2>&1 )
cat gpurun_out/r2m_tests.log | tail -3; cat gpurun_out/r2m_exp_knobs.log | cut -c1-330
ls -la gpurun_out | tail -5
