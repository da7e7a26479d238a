set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --steps 3 --warmup 2 > gpurun_out/bench_pipe_3g.json 2> gpurun_out/bench_pipe_3g.err; tail -2 gpurun_out/bench_pipe_3g.err; cat gpurun_out/bench_pipe_3g.json
ncu --set full --clock-control none --import-source on -k regex:bsw_thread_kernel -s 102 -c 10 -o gpurun_out/prof_bsw_r1g python bench.py --steps 1 --warmup 2 > gpurun_out/ncu_full_bsw.log 2>&1
ls -la gpurun_out
