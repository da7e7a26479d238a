#!/bin/bash
# Round 2, GPU call W (2 GPUs): the final tree's sharded bench and reference arm as the driver launches them at N = 2.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2> gpurun_out/r2w_bench_ref_2gpu.err | tail -1 ) > gpurun_out/r2w_bench_reference_arm_2gpu.json
( timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 3 --warmup 3 2> gpurun_out/r2w_bench_2gpu.err | tail -1 ) > gpurun_out/r2w_bench_3gbp_2gpu.json
python - <<'PY'
import json
a=json.loads(open('gpurun_out/r2w_bench_3gbp_2gpu.json').read()); r=json.loads(open('gpurun_out/r2w_bench_reference_arm_2gpu.json').read())
print('same metric/config/unit:', a['metric']==r['metric'], a['config']==r['config'], a['unit']==r['unit'], 'n_gpus', a['n_gpus'], r['n_gpus'], 'value', a['value'], 'e2e', a['e2e']['value'], 'ref', r['value'], a.get('sharding',{}).get('startup_s_rank0'))
PY
tail -c 400 gpurun_out/r2w_bench_2gpu.err
