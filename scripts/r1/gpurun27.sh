set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1q.json 2> gpurun_out/bench_r1q.err; tail -2 gpurun_out/bench_r1q.err; cat gpurun_out/bench_r1q.json
timeout 200 python bench.py --workload cigar --steps 2 --warmup 1 > gpurun_out/bench_cigar_r1q.json 2> gpurun_out/bench_cigar_r1q.err; tail -3 gpurun_out/bench_cigar_r1q.err; cat gpurun_out/bench_cigar_r1q.json
timeout 200 python scripts/exp_knobs.py /tmp/bm2_bench_pipe_3000_500000 3 > gpurun_out/exp_knobs_r1q.log 2>&1; cut -c1-330 gpurun_out/exp_knobs_r1q.log
