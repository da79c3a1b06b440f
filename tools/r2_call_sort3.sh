#!/bin/bash
# step time of the single-GPU pipeline with the wide plan vs the library plan (same box, back to back)
mkdir -p gpurun_out
for s in own cub own cub; do
  ESR_PLAN_SORT=$s timeout 200 python bench.py --no-cpu --no-uniform --no-inbatch --no-table-100m > gpurun_out/sortbench_$s.json 2> gpurun_out/sortbench_$s.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/sortbench_$s.json').read().strip().splitlines()[-1])
print('$s', 'value %.3f G  ms/step %.4f  e2e %.3f G' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9))
PY
done
