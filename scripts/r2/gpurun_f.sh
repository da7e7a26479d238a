#!/bin/bash
# Round 2, GPU call F (2 GPUs): the sharded product path under torchrun: index read once on rank 0 + NCCL broadcast, different reads per rank.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
( timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2> gpurun_out/r2f_bench_2gpu.err | tail -1 ) > gpurun_out/r2f_bench_3gbp_2gpu.json
( timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 2> gpurun_out/r2f_bench_1gpu.err | tail -1 ) > gpurun_out/r2f_bench_3gbp_1gpu.json
tail -c 600 gpurun_out/r2f_bench_2gpu.err
ls -la gpurun_out | tail -5
