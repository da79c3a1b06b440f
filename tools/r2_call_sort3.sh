#!/bin/bash
# plan builders: parity of all three, standalone time, and the step with cub + fused head pass vs cub + round-1 head kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_glove.py tests/test_gpu_virtual_peers.py -x -q -m gpu -k "plan or routed" 2>&1 | tail -5 | tee gpurun_out/sort_tests.log
timeout 120 python tools/prof_plan.py 2>&1 | tail -3 | tee gpurun_out/sort_prof.json
for s in cub cub_split cub cub_split own; do
  ESR_PLAN_SORT=$s timeout 200 python bench.py --no-cpu --no-uniform --no-inbatch --no-table-100m > gpurun_out/sortbench_$s.json 2> gpurun_out/sortbench_$s.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/sortbench_$s.json').read().strip().splitlines()[-1])
print('$s', 'value %.3f G  ms/step %.4f  e2e %.3f G' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9))
PY
done
