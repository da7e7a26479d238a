#!/bin/bash
# Round 2, GPU call R: the round's final tree: whole GPU suite, smoke(), default bench + reference arm, a last sweep of the lane / CTA knobs.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
W=/tmp/bm2_bench_pipe_3000_500000
( timeout 1500 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2r_bench.err | tail -1 ) > gpurun_out/r2r_bench_3gbp_1gpu.json
( timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 ) > gpurun_out/r2r_tests.log 2>&1
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/r2r_smoke.log
( timeout 900 python scripts/exp_knobs.py $W 3 2>&1 | tail -14 ) > gpurun_out/r2r_exp_knobs.log
( timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/r2r_bench_ref.err | tail -1 ) > gpurun_out/r2r_bench_reference_arm.json
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2r_launches_step.csv python scripts/prof_step.py $W 2 > /tmp/ncu_r.log 2>&1 )
cat gpurun_out/r2r_tests.log | tail -3; cat gpurun_out/r2r_smoke.log | cut -c1-200; cat gpurun_out/r2r_exp_knobs.log | cut -c1-220
ls -la gpurun_out | tail -8
