set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_pipe_3g_2gpu.json 2> gpurun_out/bench_pipe_3g_2gpu.err; tail -5 gpurun_out/bench_pipe_3g_2gpu.err; cat gpurun_out/bench_pipe_3g_2gpu.json
python bench.py --steps 3 --warmup 2 > gpurun_out/bench_pipe_3g.json 2> gpurun_out/bench_pipe_3g.err; tail -2 gpurun_out/bench_pipe_3g.err; cat gpurun_out/bench_pipe_3g.json
ls -la gpurun_out
