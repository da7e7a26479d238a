set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --steps 3 --warmup 2 > gpurun_out/bench_pipe_3g.json 2> gpurun_out/bench_pipe_3g.err; tail -2 gpurun_out/bench_pipe_3g.err; cat gpurun_out/bench_pipe_3g.json
python scripts/exp_multi_ctx.py /tmp/bm2_bench_pipe_3000_500000 4 > gpurun_out/exp_multi_ctx.log 2>&1; cat gpurun_out/exp_multi_ctx.log | tail -8
ls -la gpurun_out
