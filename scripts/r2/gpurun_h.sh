#!/bin/bash
# Round 2, GPU call H: staggering the sub-batch lanes with the SMEM token (knob sweep), smoke(), long reads with REDUX reductions, sam bench after the reverts.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
W=/tmp/bm2_bench_pipe_3000_500000
( timeout 1200 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2h_bench.err | tail -1 ) > gpurun_out/r2h_bench_3gbp_1gpu.json
( timeout 900 python scripts/exp_knobs.py $W 3 2>&1 | tail -14 ) > gpurun_out/r2h_exp_knobs.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/r2h_smoke.log
( timeout 900 python bench.py --workload longread --long-reads 512 --long-sample 512 --steps 1 --warmup 1 2> gpurun_out/r2h_bench_long.err | tail -1 ) > gpurun_out/r2h_bench_long.json
( timeout 600 python -m pytest tests/test_bsw_gpu.py tests/test_longreads_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 ) > gpurun_out/r2h_tests.log 2>&1
cat gpurun_out/r2h_smoke.log; cat gpurun_out/r2h_tests.log | tail -2
ls -la gpurun_out | tail -6
