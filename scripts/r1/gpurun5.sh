set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_pipe.json 2> gpurun_out/bench_pipe.err; tail -3 gpurun_out/bench_pipe.err; cat gpurun_out/bench_pipe.json
ncu --set full --clock-control none --import-source on -k regex:bsw_thread_kernel -s 64 -c 16 -o gpurun_out/prof_bsw_r1b python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full_bsw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:smem_kernel -s 4 -c 1 -o gpurun_out/prof_smem_r1b python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full_smem.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:chain_kernel|tail_kernel" -s 8 -c 2 -o gpurun_out/prof_chain_tail_r1b python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full_chain.log 2>&1
ls -la gpurun_out
