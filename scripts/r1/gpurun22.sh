set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err; tail -3 gpurun_out/bench_r1m.err; cat gpurun_out/bench_r1m.json
timeout 420 python scripts/exp_knobs.py /tmp/bm2_bench_pipe_3000_500000 3 > gpurun_out/exp_knobs_r1m.log 2>&1; cat gpurun_out/exp_knobs_r1m.log
timeout 420 ncu --set full --clock-control none --import-source on -k regex:bsw_col2_kernel -c 6 -o gpurun_out/prof_bsw_col2_r1m python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full_col2.log 2>&1
tail -2 gpurun_out/ncu_full_col2.log
ls -la gpurun_out
