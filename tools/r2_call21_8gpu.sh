#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_sharded.py -q -x -k "sharded_matches and routed" 2>&1 | tail -n 3
show() { python - $1 <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" N=%d value %.3f G  ms %.4f  e2e %.3f G  parity %s" % (d["n_gpus"], d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d.get("parity_check",{}).get("ok")))
    for k in ("table_100m","inbatch_sharded"):
        if k in d: print("  ", k, json.dumps(d[k])[:500])
except Exception as e:
    print(" parse error", e); print(open(sys.argv[1].replace(".json",".err")).read()[-3000:])
PY
}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 8 --steps 40 --warmup 10 > gpurun_out/r2c21_bench_8gpu.json 2> gpurun_out/r2c21_bench_8gpu.err
show gpurun_out/r2c21_bench_8gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus 4 --steps 40 --warmup 10 --no-inbatch > gpurun_out/r2c21_bench_4gpu.json 2> gpurun_out/r2c21_bench_4gpu.err
show gpurun_out/r2c21_bench_4gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 \
    bench.py --gpus 8 --steps 20 --warmup 5 --no-inbatch --no-table-100m > gpurun_out/r2c21_bench_8gpu_20.json 2> gpurun_out/r2c21_bench_8gpu_20.err
show gpurun_out/r2c21_bench_8gpu_20.json
