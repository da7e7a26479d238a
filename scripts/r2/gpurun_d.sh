#!/bin/bash
# Round 2, GPU call D: whole suite (staged rescue default, FASTQ / SAM-text seams, sharded start-up), bench with the register band shrink and
# per-class CTA sizes, long reads with the warp-shared seed alignments, FASTQ -> SAM end to end.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
( timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/r2d_tests_gpu.log 2>&1
( timeout 1200 python bench.py --steps 5 --warmup 3 2> gpurun_out/r2d_bench.err | tail -1 ) > gpurun_out/r2d_bench_3gbp_1gpu.json
( timeout 900 python bench.py --workload longread --long-reads 512 --long-sample 512 --steps 1 --warmup 1 2> gpurun_out/r2d_bench_long.err | tail -1 ) > gpurun_out/r2d_bench_long.json
( timeout 900 python bench.py --workload fastq2sam --steps 2 --warmup 1 2> gpurun_out/r2d_bench_fastq2sam.err | tail -1 ) > gpurun_out/r2d_bench_fastq2sam.json
tail -c 300 gpurun_out/r2d_bench_fastq2sam.err; tail -c 300 gpurun_out/r2d_bench_long.err; cat gpurun_out/r2d_tests_gpu.log | tail -4
ls -la gpurun_out | tail -8
