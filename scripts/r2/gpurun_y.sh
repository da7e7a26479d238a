#!/bin/bash
# Round 2, GPU call Y: smoke() and the extension-kernel tests on the library as rebuilt from a clean tree.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
( timeout 240 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/r2y_smoke.log
( timeout 200 python -m pytest tests/test_bsw_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2 ) > gpurun_out/r2y_tests.log 2>&1
cat gpurun_out/r2y_smoke.log gpurun_out/r2y_tests.log | cut -c1-250
