set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_bsw.json 2> gpurun_out/bench_bsw.err; tail -2 gpurun_out/bench_bsw.err; cat gpurun_out/bench_bsw.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bsw_thread_kernel -s 16 -c 8 -o gpurun_out/prof_bsw_r1a python bench.py --steps 1 --warmup 1 --bsw-jobs 1000000 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
