#!/bin/bash
mkdir -p gpurun_out
N=${N:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 40 --warmup 10 > gpurun_out/r2c13_bench_${N}gpu.json 2> gpurun_out/r2c13_bench_${N}gpu.err
python - gpurun_out/r2c13_bench_${N}gpu.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" value %.3f G  ms %.4f  e2e %.3f G  parity %s" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d.get("parity_check",{}).get("ok")))
    print("  nvlink", json.dumps(d.get("roofline",{}).get("nvlink")))
    for k in ("table_100m","inbatch_sharded"):
        if k in d: print("  ", k, json.dumps(d[k])[:900])
except Exception as e:
    print(" parse error", e); print(open(sys.argv[1].replace(".json",".err")).read()[-3000:])
PY
timeout 200 python -m torch.distributed.run --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 tools/prof_routed.py 2>/dev/null | tail -n 1 | tee gpurun_out/r2c13_prof_routed_w${N}.json
V=100000000 timeout 300 python -m torch.distributed.run --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 tools/prof_routed.py 2>/dev/null | tail -n 1 | tee gpurun_out/r2c13_prof_routed_100m_w${N}.json
