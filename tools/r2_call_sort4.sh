#!/bin/bash
# 2 GPUs: owner-routed step with the hand-written plan vs the cub control
mkdir -p gpurun_out
for s in own cub own cub; do
  ESR_PLAN_SORT=$s timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 --no-table-100m --no-inbatch > gpurun_out/sortbench2_$s.json 2> gpurun_out/sortbench2_$s.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/sortbench2_$s.json').read().strip().splitlines()[-1])
    print('$s', 'value %.3f G  ms/step %.4f  e2e %.3f G  parity %s' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d.get('parity_check',{}).get('ok')))
except Exception as e:
    print('$s failed', e); print(open('gpurun_out/sortbench2_$s.err').read()[-1500:])
PY
done
