#!/bin/bash
# 1 GPU: new parity tests (owner-routed virtual ranks, n_valid plumbing, interleave variant, bench shape), bench A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_virtual_peers.py tests/test_gpu_glove.py tests/test_gpu_sharded.py -x -q > gpurun_out/r2c4_tests.log 2>&1
tail -n 15 gpurun_out/r2c4_tests.log
timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu --no-inbatch > gpurun_out/r2c4_bench_default.json 2> gpurun_out/r2c4_bench_default.err
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --no-inbatch --no-table-100m --kernel interleave > gpurun_out/r2c4_bench_interleave.json 2> gpurun_out/r2c4_bench_interleave.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-inbatch --no-table-100m > gpurun_out/r2c4_bench_default_20.json 2> gpurun_out/r2c4_bench_default_20.err
for f in gpurun_out/r2c4_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" value %.3f G  e2e %.3f G  ms/step %.4f  frac zipf %.3f (%.1f us)  frac unif %s" % (d["value"]/1e9, d["e2e"]["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"]*1e3, d.get("roofline_uniform",{}).get("frac")))
    if "table_100m" in d: print(" table_100m", json.dumps(d["table_100m"])[:600])
except Exception as e:
    print(" parse error", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
