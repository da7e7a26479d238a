set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --steps 3 --warmup 2 > gpurun_out/bench_pipe_3g.json 2> gpurun_out/bench_pipe_3g.err; tail -2 gpurun_out/bench_pipe_3g.err; cat gpurun_out/bench_pipe_3g.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 600 --csv --log-file gpurun_out/launches_r1e_3g.csv python bench.py --steps 1 --warmup 2 > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out
