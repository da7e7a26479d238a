set -x
mkdir -p gpurun_out
python scripts/debug_sub_batches.py > gpurun_out/debug_sub.log 2>&1; cat gpurun_out/debug_sub.log | tail -40
