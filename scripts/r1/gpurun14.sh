set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --steps 3 --warmup 2 > gpurun_out/bench_pipe_3g.json 2> gpurun_out/bench_pipe_3g.err; tail -2 gpurun_out/bench_pipe_3g.err; cat gpurun_out/bench_pipe_3g.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:^(smem_|bsw_|chain_|tail_|sa_|ext_|fold_|right_|gather_|regs_|slot_|mark_|work_|read_)' --csv --log-file gpurun_out/r1h_traffic.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_traffic.log 2>&1
tail -3 gpurun_out/ncu_traffic.log
ls -la gpurun_out
