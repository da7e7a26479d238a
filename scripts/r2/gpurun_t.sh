#!/bin/bash
# Round 2, GPU call T: the warp-per-read tail kernel compiled for 4 / 6 / 8 resident CTAs per SM (128 / 80 / 64 registers): occupancy against spills.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
W=/tmp/bm2_bench_pipe_3000_500000
( timeout 1200 python scripts/exp_knobs.py $W 3 2>&1 | grep -v index_build | tail -12 ) > gpurun_out/r2t_exp_knobs.log 2> gpurun_out/r2t_exp_knobs.err
( timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_zz_tandem_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 ) > gpurun_out/r2t_tests.log 2>&1
cat gpurun_out/r2t_tests.log | tail -2; cat gpurun_out/r2t_exp_knobs.log | cut -c1-260
