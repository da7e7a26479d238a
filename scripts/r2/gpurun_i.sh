#!/bin/bash
# Round 2, GPU call I: lane-skew sweep, SAM / CIGAR benches with the load-hoisted DP, FASTQ -> SAM, source-level profile of the tail and chain kernels.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
W=/tmp/bm2_bench_pipe_3000_500000
( timeout 1200 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2i_bench.err | tail -1 ) > gpurun_out/r2i_bench_3gbp_1gpu.json
( timeout 900 python scripts/exp_knobs.py $W 3 2>&1 | tail -12 ) > gpurun_out/r2i_exp_knobs.log
( timeout 600 python -m pytest tests/test_cigar_gpu.py tests/test_zz_sam_gpu.py tests/test_zzz_sam_staged_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 ) > gpurun_out/r2i_tests.log 2>&1
( timeout 900 python bench.py --workload sam --steps 2 --warmup 1 2> gpurun_out/r2i_bench_sam.err | tail -1 ) > gpurun_out/r2i_bench_sam.json
( timeout 900 python bench.py --workload cigar --steps 2 --warmup 1 2> gpurun_out/r2i_bench_cigar.err | tail -1 ) > gpurun_out/r2i_bench_cigar.json
( timeout 900 python bench.py --workload fastq2sam --steps 2 --warmup 1 2> gpurun_out/r2i_bench_fastq2sam.err | tail -1 ) > gpurun_out/r2i_bench_fastq2sam.json
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tail_kernel|chain_kernel' -s 4 -c 4 -o /tmp/r2i_ct python scripts/prof_step.py $W 2 > /tmp/ncu_ct.log 2>&1 ;
  [ -f /tmp/r2i_ct.ncu-rep ] && ncu -i /tmp/r2i_ct.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2i_src_chain_tail.csv.gz ) > gpurun_out/r2i_ncu_ct.log 2>&1
cat gpurun_out/r2i_tests.log | tail -2; tail -c 200 gpurun_out/r2i_bench_cigar.err
ls -la gpurun_out | tail -8
