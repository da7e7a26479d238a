set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --ref-mbp 3000 --steps 3 --warmup 2 > gpurun_out/bench_pipe_3g.json 2> gpurun_out/bench_pipe_3g.err; tail -2 gpurun_out/bench_pipe_3g.err; cat gpurun_out/bench_pipe_3g.json
ncu --set full --clock-control none --import-source on -k regex:smem_kernel -s 3 -c 1 -o gpurun_out/prof_smem_r1c python bench.py --ref-mbp 3000 --steps 1 --warmup 2 > gpurun_out/ncu_full_smem.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bsw_thread_kernel -s 102 -c 34 -o gpurun_out/prof_bsw_r1c python bench.py --ref-mbp 3000 --steps 1 --warmup 2 > gpurun_out/ncu_full_bsw.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/launches_r1c_3g.csv python bench.py --ref-mbp 3000 --steps 1 --warmup 2 > gpurun_out/ncu_bench.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_pipe.json 2> gpurun_out/bench_pipe.err; cat gpurun_out/bench_pipe.json
ls -la gpurun_out
