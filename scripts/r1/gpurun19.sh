set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --steps 3 --warmup 2 > gpurun_out/bench_pipe_3g.json 2> gpurun_out/bench_pipe_3g.err; tail -2 gpurun_out/bench_pipe_3g.err; cat gpurun_out/bench_pipe_3g.json
python bench.py --steps 3 --warmup 2 --sub-batches 2 > gpurun_out/bench_pipe_3g_sb2.json 2> gpurun_out/bench_pipe_3g_sb2.err; cat gpurun_out/bench_pipe_3g_sb2.json
python bench.py --steps 3 --warmup 2 --sub-batches 8 > gpurun_out/bench_pipe_3g_sb8.json 2> gpurun_out/bench_pipe_3g_sb8.err; cat gpurun_out/bench_pipe_3g_sb8.json
ls -la gpurun_out
