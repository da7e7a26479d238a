#!/bin/bash
# Round 2, GPU call B: the rewritten column-pair extension kernel (uniform control flow): parity tests, unsplit stage times, full bench,
# ncu launch list of one unsplit step, full captures of the extension kernels and of the chain / tail kernels.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
W=/tmp/bm2_bench_pipe_3000_500000
( timeout 600 python -m pytest tests/test_bsw_gpu.py tests/test_pipeline_gpu.py tests/test_dropin_sam_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/r2b_tests_bsw.log 2>&1
( timeout 1200 python bench.py --steps 5 --warmup 3 2> gpurun_out/r2b_bench.err | tail -1 ) > gpurun_out/r2b_bench_3gbp_1gpu.json
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 115 -c 115 --csv --log-file gpurun_out/r2b_launches_step.csv \
    python scripts/prof_step.py $W 2 > gpurun_out/r2b_prof_step.log 2>&1 )
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:bsw_col2 -s 12 -c 12 -o /tmp/r2b_bsw python scripts/prof_step.py $W 2 > /tmp/ncu_bsw.log 2>&1 ;
  [ -f /tmp/r2b_bsw.ncu-rep ] && python scripts/ncu_summary.py /tmp/r2b_bsw.ncu-rep gpurun_out/r2b_bsw_col2.md 'bsw_col2_kernel (round 2: uniform control flow), the 12 launches of one unsplit 1 M-read step' ) > gpurun_out/r2b_ncu_bsw.log 2>&1
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tail_kernel|chain_kernel|ext_build' -s 5 -c 5 -o /tmp/r2b_ct python scripts/prof_step.py $W 2 > /tmp/ncu_ct.log 2>&1 ;
  [ -f /tmp/r2b_ct.ncu-rep ] && python scripts/ncu_summary.py /tmp/r2b_ct.ncu-rep gpurun_out/r2b_chain_tail.md 'chain_kernel / ext_build_kernel / tail_kernel of one unsplit 1 M-read step' ) > gpurun_out/r2b_ncu_ct.log 2>&1
tail -c 300 gpurun_out/r2b_bench.err; cat gpurun_out/r2b_tests_bsw.log | tail -3
ls -la gpurun_out | tail -12
