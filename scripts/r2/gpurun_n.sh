#!/bin/bash
# Round 2, GPU call N: A/B of the warp-cooperative heavy tail and of the register-edge band shrink, then the whole GPU suite, smoke(), the default
# bench with its reference arm, and the benches of the other workloads with the round's final code.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
W=/tmp/bm2_bench_pipe_3000_500000
( timeout 1500 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2n_bench.err | tail -1 ) > gpurun_out/r2n_bench_3gbp_1gpu.json
( timeout 900 python scripts/exp_knobs.py $W 3 2>&1 | tail -16 ) > gpurun_out/r2n_exp_knobs.log
( timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 ) > gpurun_out/r2n_tests.log 2>&1
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/r2n_smoke.log
( timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/r2n_bench_ref.err | tail -1 ) > gpurun_out/r2n_bench_reference_arm.json
( timeout 900 python bench.py --workload fastq2sam --steps 2 --warmup 1 2> gpurun_out/r2n_bench_fastq2sam.err | tail -1 ) > gpurun_out/r2n_bench_fastq2sam.json
( timeout 900 python bench.py --workload sam --steps 2 --warmup 1 2> gpurun_out/r2n_bench_sam.err | tail -1 ) > gpurun_out/r2n_bench_sam.json
( timeout 900 python bench.py --workload longread --steps 2 --warmup 1 2> gpurun_out/r2n_bench_long.err | tail -1 ) > gpurun_out/r2n_bench_long.json
cat gpurun_out/r2n_tests.log | tail -3; cat gpurun_out/r2n_smoke.log; cat gpurun_out/r2n_exp_knobs.log | cut -c1-330
ls -la gpurun_out | tail -10
