#!/bin/bash
# after AUTO = wide: GloVe / model GPU tests, smoke, and the bench line (with the 100M-row table sub-record)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_glove.py tests/test_gpu_models.py tests/test_gpu_inbatch.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_final2_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --no-cpu --no-inbatch > gpurun_out/r2_final2_bench.json 2> gpurun_out/r2_final2_bench.err; tail -c 400 gpurun_out/r2_final2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final2_bench.json').read().strip().splitlines()[-1])
print('value %.3f G  ms/step %.4f  e2e %.3f G  frac %.3f  unif %s launches %d' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['roofline']['frac'], d.get('roofline_uniform',{}).get('frac'), d['gpu_launches']))
print('table_100m', d.get('table_100m',{}).get('pairs_per_s'), d.get('table_100m',{}).get('ms_per_step'))
PY
