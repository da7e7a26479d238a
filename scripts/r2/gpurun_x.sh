#!/bin/bash
# Round 2, GPU call X: bm2_mem with gzip input against the reference on the same .gz files (and the other FASTQ -> SAM tests).
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
( timeout 600 python -m pytest tests/test_zz_fastq_sam_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 ) > gpurun_out/r2x_tests.log 2>&1
cat gpurun_out/r2x_tests.log
