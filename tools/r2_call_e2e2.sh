#!/bin/bash
# 2 GPUs: sharded trainer tests (routed variant) + owner-routed bench with the finish-kernel loss mirror
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu -k "routed" 2>&1 | tail -5 | tee gpurun_out/e2e2_tests.log
for s in a b; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 --no-table-100m --no-inbatch > gpurun_out/e2e2_$s.json 2> gpurun_out/e2e2_$s.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/e2e2_$s.json').read().strip().splitlines()[-1])
    print('run $s', 'value %.3f G  ms/step %.4f  e2e %.3f G (%.4f ms)  parity %s' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], d.get('parity_check',{}).get('ok')))
except Exception as e:
    print('$s failed', e); print(open('gpurun_out/e2e2_$s.err').read()[-2500:])
PY
done
