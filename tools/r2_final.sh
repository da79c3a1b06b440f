#!/bin/bash
# final validation of round 2 on one B200: full GPU suite, smoke, default bench line, sanitizer on the late kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2_final_smoke.log
timeout 600 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -c 600 gpurun_out/r2_final_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench.json').read().strip().splitlines()[-1])
print('value %.3f G  ms/step %.4f  e2e %.3f G  frac %.3f  unif %s' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['roofline']['frac'], d.get('roofline_uniform',{}).get('frac')))
print('table_100m', d.get('table_100m',{}).get('value'), 'inbatch', {k: round(v['ms_per_step'],4) for k,v in d.get('other_workloads',{}).items() if isinstance(v, dict) and 'ms_per_step' in v})
PY
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_new_kernels.py 2>&1 | tail -4 | tee gpurun_out/r2_final_sanitizer.txt
