#!/bin/bash
# last 2 GPU-minutes of the round: the two test files added after the budget ran out (chain B-tree shape on the device; seam 4 first run)
mkdir -p gpurun_out
( timeout 50 python -m pytest tests/test_zz_tandem_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/zz_tandem.log 2>&1
( timeout 60 python -m pytest tests/test_zz_sam_gpu.py -q --runxfail -p no:cacheprovider 2>&1 | tail -40 ) > gpurun_out/zz_sam.log 2>&1
echo done
