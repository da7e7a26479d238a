#!/bin/bash
# Round 2, GPU call V: the two arms of the default bench exactly as the driver runs them, after the config object became common to both lines.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
( timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 2> gpurun_out/r2v_bench_ref.err | tail -1 ) > gpurun_out/r2v_bench_reference_arm.json
( timeout 1500 python bench.py --gpus 1 --steps 3 --warmup 3 2> gpurun_out/r2v_bench.err | tail -1 ) > gpurun_out/r2v_bench_3gbp_1gpu.json
python - <<'PY'
import json
a=json.loads(open('gpurun_out/r2v_bench_3gbp_1gpu.json').read()); r=json.loads(open('gpurun_out/r2v_bench_reference_arm.json').read())
print('same metric/config/unit:', a['metric']==r['metric'], a['config']==r['config'], a['unit']==r['unit'], 'value', a['value'], 'e2e', a['e2e']['value'], 'ref', r['value'], 'e2e ratio', a['e2e']['value']/r['value'])
PY
tail -c 300 gpurun_out/r2v_bench.err; tail -c 300 gpurun_out/r2v_bench_ref.err
