#!/bin/bash
# Round 2, GPU call P: waves with a copy stream and two output buffers per lane: parity, end-to-end A/B by (lanes, waves).
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
W=/tmp/bm2_bench_pipe_3000_500000
( timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_shard_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 ) > gpurun_out/r2p_tests.log 2>&1
( timeout 1500 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2p_bench.err | tail -1 ) > gpurun_out/r2p_bench_3gbp_1gpu.json
( timeout 900 python scripts/exp_knobs.py $W 3 2>&1 | tail -14 ) > gpurun_out/r2p_exp_knobs.log
cat gpurun_out/r2p_tests.log | tail -3; cat gpurun_out/r2p_exp_knobs.log | cut -c1-200
ls -la gpurun_out | tail -5
