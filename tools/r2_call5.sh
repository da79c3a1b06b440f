#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_glove.py tests/test_ref_golden.py -m gpu -x -q > gpurun_out/r2c5_tests.log 2>&1
tail -n 25 gpurun_out/r2c5_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout 500 python bench.py --steps 100 --warmup 10 --no-cpu --no-table-100m > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-inbatch --no-table-100m --no-uniform > gpurun_out/r2c5_bench20.json 2> gpurun_out/r2c5_bench20.err
for f in gpurun_out/r2c5_bench.json gpurun_out/r2c5_bench20.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" value %.3f G  e2e %.3f G  ms/step %.4f  e2e ms %.4f frac zipf %.3f (%.1f us)  frac unif %s" % (d["value"]/1e9, d["e2e"]["value"]/1e9, d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"]*1e3, d.get("roofline_uniform",{}).get("frac")))
    if "retrieval" in d: print(" retrieval", json.dumps(d["retrieval"])[:1500])
except Exception as e:
    print(" parse error", e); print(open(sys.argv[1].replace(".json",".err")).read()[-2500:])
PY
done
