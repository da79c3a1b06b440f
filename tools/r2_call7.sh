#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_models.py tests/test_gpu_glove.py -q -x > gpurun_out/r2c7_tests.log 2>&1; tail -n 12 gpurun_out/r2c7_tests.log
for T in OwnerRoutedGloveTrainer PeerShardedGloveTrainer; do
  TRAINER=$T timeout 200 python -m torch.distributed.run --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29650 tools/prof_routed.py 2>/dev/null | tail -n 1 | tee gpurun_out/r2c7_prof_${T}_w1.json
done
timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu --no-table-100m > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-inbatch --no-table-100m --no-uniform > gpurun_out/r2c7_bench20.json 2> gpurun_out/r2c7_bench20.err
for f in gpurun_out/r2c7_bench.json gpurun_out/r2c7_bench20.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" value %.3f G  e2e %.3f G  ms/step %.4f  e2e ms %.4f frac zipf %.3f (%.1f us)  frac unif %s" % (d["value"]/1e9, d["e2e"]["value"]/1e9, d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"]*1e3, d.get("roofline_uniform",{}).get("frac")))
    if "retrieval" in d: print(" retrieval", json.dumps(d["retrieval"])[:1200])
    if "other_workloads" in d: print(" other", json.dumps(d["other_workloads"])[:3000])
except Exception as e:
    print(" parse error", e); print(open(sys.argv[1].replace(".json",".err")).read()[-2500:])
PY
done
