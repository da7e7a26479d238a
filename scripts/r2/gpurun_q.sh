#!/bin/bash
# Round 2, GPU call Q (4 GPUs of one box): the sharded bench as the driver launches it (one rank per GPU, index read once and broadcast, different
# reads per rank), the reference arm under torchrun (rank 0 alone works), the gloo-free GPU shard test.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi -L | head -8 > gpurun_out/r2q_gpus.log
( timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 3 --warmup 3 2> gpurun_out/r2q_bench_4gpu.err | tail -1 ) > gpurun_out/r2q_bench_3gbp_4gpu.json
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 2> gpurun_out/r2q_bench_ref_4gpu.err | tail -1 ) > gpurun_out/r2q_bench_reference_arm_4gpu.json
tail -c 1500 gpurun_out/r2q_bench_4gpu.err; cut -c1-600 gpurun_out/r2q_bench_3gbp_4gpu.json; cut -c1-300 gpurun_out/r2q_bench_reference_arm_4gpu.json
ls -la gpurun_out | tail -5
