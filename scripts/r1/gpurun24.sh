set -x
mkdir -p gpurun_out
W=/tmp/bm2_bench_pipe_3000_500000
timeout 600 python scripts/prof_step.py $W 1 > gpurun_out/prep.log 2>&1; tail -2 gpurun_out/prep.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r1n_launches.csv python scripts/prof_step.py $W 2 > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:bsw_col2_kernel|tail_kernel|chain_kernel|smem_bwd_kernel|smem_fwd1_kernel' -s 19 -c 19 -o /tmp/prof_step_r1n python scripts/prof_step.py $W 2 > gpurun_out/ncu_full_step.log 2>&1; tail -2 gpurun_out/ncu_full_step.log
python scripts/ncu_summary.py /tmp/prof_step_r1n.ncu-rep gpurun_out/r1n_step_kernels.md "round 1n: col2 BSW, chain, tail, smem_fwd1, smem_bwd kernels of one unsplit 1 M-read step, 3 Gbp"
ncu -i /tmp/prof_step_r1n.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r1n_step_raw.csv.gz
for k in bsw_col2_kernel tail_kernel chain_kernel; do
  ncu -i /tmp/prof_step_r1n.ncu-rep --page source --csv -k regex:$k 2>/dev/null | gzip > gpurun_out/r1n_src_$k.csv.gz
done
ls -la gpurun_out /tmp/prof_step_r1n.ncu-rep
du -sh gpurun_out
