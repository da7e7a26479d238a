#!/bin/bash
# Round 2, GPU call E: knob sweep (extension-kernel CTA size A/B, shared memory left to the SMEM kernels of the other lanes), the C++ host
# program (test + fastq2sam bench), source-level profile of the per-pair SAM kernel in staged mode, bench line with the TMA gather probe.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
W=/tmp/bm2_bench_pipe_3000_500000
( timeout 1200 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2e_bench.err | tail -1 ) > gpurun_out/r2e_bench_3gbp_1gpu.json
( timeout 900 python scripts/exp_knobs.py $W 3 2>&1 | tail -30 ) > gpurun_out/r2e_exp_knobs.log
( timeout 900 python -m pytest tests/test_zz_fastq_sam_gpu.py tests/test_longreads_gpu.py tests/test_zz_tandem_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/r2e_tests_fastq.log 2>&1
( timeout 900 python bench.py --workload longread --long-reads 512 --long-sample 512 --steps 1 --warmup 1 2> gpurun_out/r2e_bench_long.err | tail -1 ) > gpurun_out/r2e_bench_long.json
( timeout 900 python bench.py --workload fastq2sam --steps 2 --warmup 1 2> gpurun_out/r2e_bench_fastq2sam.err | tail -1 ) > gpurun_out/r2e_bench_fastq2sam.json
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^.*sam_kernel' -s 1 -c 1 -o /tmp/r2e_sam python bench.py --workload sam --steps 1 --warmup 0 > /tmp/ncu_sam.log 2>&1 ;
  [ -f /tmp/r2e_sam.ncu-rep ] && python scripts/ncu_summary.py /tmp/r2e_sam.ncu-rep gpurun_out/r2e_sam_kernel.md 'sam_kernel (per-pair logic of the SAM stage)' &&
  ncu -i /tmp/r2e_sam.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2e_src_sam_kernel.csv.gz ) > gpurun_out/r2e_ncu_sam.log 2>&1
tail -3 /tmp/ncu_sam.log; tail -c 300 gpurun_out/r2e_bench_fastq2sam.err; cat gpurun_out/r2e_tests_fastq.log | tail -3
ls -la gpurun_out | tail -8
