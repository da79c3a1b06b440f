#!/bin/bash
mkdir -p gpurun_out
N=${N:-2}
timeout 100 python -m pytest tests/test_gpu_sharded.py -q -x -k "route_plan" 2>&1 | tail -n 3
timeout 300 python -m pytest tests/test_gpu_sharded.py -q -x -k "sharded_matches and routed" 2>&1 | tail -n 3
run() {  # name, extra flags
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $N --steps 100 --warmup 10 $2 > gpurun_out/r2c11_bench_${N}gpu_$1.json 2> gpurun_out/r2c11_bench_${N}gpu_$1.err
  echo "== $1"; python - gpurun_out/r2c11_bench_${N}gpu_$1.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" value %.3f G  ms %.4f  e2e %.3f G  parity %s" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d.get("parity_check",{}).get("ok")))
    for k in ("table_100m","inbatch_sharded"):
        if k in d: print("  ", k, json.dumps(d[k])[:700])
except Exception as e:
    print(" parse error", e); print(open(sys.argv[1].replace(".json",".err")).read()[-2000:])
PY
}
run routed "--no-inbatch --no-table-100m"
run routed_fullgrid "--no-inbatch --no-table-100m --row-blocks 0"
run routed_all ""
timeout 200 python -m torch.distributed.run --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 tools/prof_routed.py 2>/dev/null | tail -n 1 | tee gpurun_out/r2c11_prof_routed_w${N}.json
timeout 200 python tools/prof_e2e.py 2>&1 | head -n 1 | cut -c1-900
