#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -q -m gpu -x ) > gpurun_out/r2c17_gpu_suite.log 2>&1; tail -n 8 gpurun_out/r2c17_gpu_suite.log
( time timeout 900 python bench.py ) > gpurun_out/r2c17_bench_default.json 2> gpurun_out/r2c17_bench_default.err
tail -n 4 gpurun_out/r2c17_bench_default.err
python - gpurun_out/r2c17_bench_default.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(" value %.3f G  e2e %.3f G  ms/step %.4f  e2e ms %.4f frac zipf %.3f (%.1f us)  frac unif %s" % (d["value"]/1e9, d["e2e"]["value"]/1e9, d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"]*1e3, d.get("roofline_uniform",{}).get("frac")))
    for k in ("retrieval","other_workloads","table_100m","cpu_baseline"):
        if k in d: print(" ",k, json.dumps(d[k])[:2500])
except Exception as e:
    print(" parse error", e); print(open(sys.argv[1].replace(".json",".err")).read()[-2500:])
PY
