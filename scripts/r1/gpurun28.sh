set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 330 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1r.json 2> gpurun_out/bench_r1r.err; tail -2 gpurun_out/bench_r1r.err; cat gpurun_out/bench_r1r.json
