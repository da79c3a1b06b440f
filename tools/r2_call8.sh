#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_virtual_peers.py tests/test_gpu_sharded.py -q -x > gpurun_out/r2c8_tests.log 2>&1; tail -n 12 gpurun_out/r2c8_tests.log
TRAINER=OwnerRoutedGloveTrainer timeout 200 python -m torch.distributed.run --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29650 tools/prof_routed.py 2>/dev/null | tail -n 1 | tee gpurun_out/r2c8_prof_routed_w1.json
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --no-inbatch --no-table-100m --no-uniform > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-inbatch --no-table-100m --no-uniform > gpurun_out/r2c8_bench20.json 2> gpurun_out/r2c8_bench20.err
for f in gpurun_out/r2c8_bench.json gpurun_out/r2c8_bench20.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" value %.3f G  e2e %.3f G  ms/step %.4f  e2e ms %.4f frac zipf %.3f (%.1f us)" % (d["value"]/1e9, d["e2e"]["value"]/1e9, d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"]*1e3))
except Exception as e:
    print(" parse error", e); print(open(sys.argv[1].replace(".json",".err")).read()[-2500:])
PY
done
timeout 200 ncu --set full --clock-control none -k regex:k_topk_scan -s 2 -c 1 --csv --page raw --log-file gpurun_out/r2c8_ncu_topk.csv python tools/prof_topk.py > gpurun_out/r2c8_ncu_topk.log 2>&1
tail -n 2 gpurun_out/r2c8_ncu_topk.log
