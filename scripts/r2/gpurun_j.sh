#!/bin/bash
# Round 2, GPU call J: the CIGAR DP with its band in registers: parity tests, SAM / CIGAR / FASTQ->SAM benches, source profile of the staged per-pair kernel.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
W=/tmp/bm2_bench_pipe_3000_500000
( timeout 900 python -m pytest tests/test_cigar_gpu.py tests/test_zz_sam_gpu.py tests/test_zzz_sam_staged_gpu.py tests/test_zz_fastq_sam_gpu.py tests/test_dropin_sam_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 ) > gpurun_out/r2j_tests.log 2>&1
( timeout 1200 python bench.py --workload sam --steps 2 --warmup 1 2> gpurun_out/r2j_bench_sam.err | tail -1 ) > gpurun_out/r2j_bench_sam.json
( timeout 900 python bench.py --workload cigar --steps 2 --warmup 1 2> gpurun_out/r2j_bench_cigar.err | tail -1 ) > gpurun_out/r2j_bench_cigar.json
( timeout 900 python bench.py --workload fastq2sam --steps 2 --warmup 1 2> gpurun_out/r2j_bench_fastq2sam.err | tail -1 ) > gpurun_out/r2j_bench_fastq2sam.json
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^.*sam_kernel' -s 3 -c 1 -o /tmp/r2j_sam python bench.py --workload sam --steps 1 --warmup 0 > /tmp/ncu_sam.log 2>&1 ;
  [ -f /tmp/r2j_sam.ncu-rep ] && python scripts/ncu_summary.py /tmp/r2j_sam.ncu-rep gpurun_out/r2j_sam_kernel_staged.md 'sam_kernel, staged rescue, CIGAR DP with the band in registers' &&
  ncu -i /tmp/r2j_sam.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2j_src_sam_kernel.csv.gz ) > gpurun_out/r2j_ncu_sam.log 2>&1
cat gpurun_out/r2j_tests.log | tail -3
ls -la gpurun_out | tail -6
