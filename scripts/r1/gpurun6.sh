set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --ref-mbp 3000 --steps 3 --warmup 2 > gpurun_out/bench_pipe_3g.json 2> gpurun_out/bench_pipe_3g.err; tail -30 gpurun_out/bench_pipe_3g.err; cat gpurun_out/bench_pipe_3g.json
python bench.py --impl reference --ref-mbp 3000 --steps 1 --warmup 1 > gpurun_out/bench_pipe_3g_ref.json 2> gpurun_out/bench_pipe_3g_ref.err; tail -3 gpurun_out/bench_pipe_3g_ref.err; cat gpurun_out/bench_pipe_3g_ref.json
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_pipe.json 2> gpurun_out/bench_pipe.err; tail -3 gpurun_out/bench_pipe.err; cat gpurun_out/bench_pipe.json
df -h /tmp | tail -1; free -g | head -2
ls -la gpurun_out
