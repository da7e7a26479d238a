set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_r1n.json 2> gpurun_out/bench_r1n.err; tail -2 gpurun_out/bench_r1n.err; cat gpurun_out/bench_r1n.json
W=/tmp/bm2_bench_pipe_3000_500000
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r1n_launches.csv python scripts/prof_step.py $W 2 > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:bsw_col2_kernel|tail_kernel|chain_kernel|smem_bwd_kernel|smem_fwd1_kernel' -s 19 -c 19 -o gpurun_out/prof_step_r1n python scripts/prof_step.py $W 2 > gpurun_out/ncu_full_step.log 2>&1; tail -2 gpurun_out/ncu_full_step.log
ls -la gpurun_out
