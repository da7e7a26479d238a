set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_pipe.json 2> gpurun_out/bench_pipe.err; tail -3 gpurun_out/bench_pipe.err; cat gpurun_out/bench_pipe.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_pipe_ref.json 2> gpurun_out/bench_pipe_ref.err; cat gpurun_out/bench_pipe_ref.json
python bench.py --workload bsw --ref-mbp 10 --pairs 50000 --steps 5 --warmup 3 > gpurun_out/bench_bsw.json 2> gpurun_out/bench_bsw.err; tail -2 gpurun_out/bench_bsw.err; cat gpurun_out/bench_bsw.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/launches_r1_pipeline.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:smem_kernel -s 3 -c 1 -o gpurun_out/prof_smem_r1 python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full_smem.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bsw_thread_kernel -s 53 -c 3 -o gpurun_out/prof_bsw_r1 python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full_bsw.log 2>&1
ls -la gpurun_out
