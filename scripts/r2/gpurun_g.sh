#!/bin/bash
# Round 2, GPU call G: SAM stage with the work-sorted pair order (A/B), FASTQ -> SAM through bm2_mem in steady state, knob A/B (band shrink,
# warp chaining threshold), source-level profile of the per-pair SAM kernel in staged mode, bench line.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
W=/tmp/bm2_bench_pipe_3000_500000
( timeout 1200 python bench.py --steps 5 --warmup 3 2> gpurun_out/r2g_bench.err | tail -1 ) > gpurun_out/r2g_bench_3gbp_1gpu.json
( timeout 900 python -m pytest tests/test_zz_sam_gpu.py tests/test_zzz_sam_staged_gpu.py tests/test_zz_fastq_sam_gpu.py tests/test_bsw_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/r2g_tests.log 2>&1
( timeout 900 python bench.py --workload sam --steps 2 --warmup 1 2> gpurun_out/r2g_bench_sam.err | tail -1 ) > gpurun_out/r2g_bench_sam.json
( BM2_SAM_ORDER=0 timeout 900 python bench.py --workload sam --steps 2 --warmup 1 2> /dev/null | tail -1 ) > gpurun_out/r2g_bench_sam_input_order.json
( timeout 900 python bench.py --workload fastq2sam --steps 2 --warmup 1 2> gpurun_out/r2g_bench_fastq2sam.err | tail -1 ) > gpurun_out/r2g_bench_fastq2sam.json
( timeout 900 python scripts/exp_knobs.py $W 3 2>&1 | tail -14 ) > gpurun_out/r2g_exp_knobs.log
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^.*sam_kernel' -s 3 -c 1 -o /tmp/r2g_sam python bench.py --workload sam --steps 1 --warmup 0 > /tmp/ncu_sam.log 2>&1 ;
  [ -f /tmp/r2g_sam.ncu-rep ] && python scripts/ncu_summary.py /tmp/r2g_sam.ncu-rep gpurun_out/r2g_sam_kernel_staged.md 'sam_kernel (per-pair logic of the SAM stage), staged rescue, work-sorted order' &&
  ncu -i /tmp/r2g_sam.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2g_src_sam_kernel.csv.gz ) > gpurun_out/r2g_ncu_sam.log 2>&1
tail -3 /tmp/ncu_sam.log; tail -c 300 gpurun_out/r2g_bench_fastq2sam.err; cat gpurun_out/r2g_tests.log | tail -3
ls -la gpurun_out | tail -8
