set -x
mkdir -p gpurun_out
W=/tmp/bm2_bench_pipe_3000_500000
timeout 600 python scripts/prof_step.py $W 1 > gpurun_out/prep.log 2>&1; tail -1 gpurun_out/prep.log
timeout 600 python scripts/exp_knobs.py $W 3 > gpurun_out/exp_knobs_r1o.log 2>&1; cat gpurun_out/exp_knobs_r1o.log | cut -c1-420
