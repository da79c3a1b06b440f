#!/bin/bash
# 2 GPUs: owner-routed sharded trainer -- parity (real NVLink, CUDA graphs) and bench vs the round-1 peer path
mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m pytest tests/test_gpu_sharded.py -q -x -k "sharded_matches" > gpurun_out/r2c6_sharded_tests.log 2>&1
tail -n 5 gpurun_out/r2c6_sharded_tests.log
run() {  # name, extra flags
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $N --steps 100 --warmup 10 $2 > gpurun_out/r2c6_bench_${N}gpu_$1.json 2> gpurun_out/r2c6_bench_${N}gpu_$1.err
  echo "== $1"; tail -n 1 gpurun_out/r2c6_bench_${N}gpu_$1.json | head -c 3000; echo; tail -n 4 gpurun_out/r2c6_bench_${N}gpu_$1.err | cut -c1-400
}
run routed ""
run peer_best "--exchange peer --fast-sync --overlap-ids --step-graphs --no-table-100m"
timeout 200 python -m pytest tests/test_gpu_models.py -q -x -k "topk or find_top_k or eval_step or find_knn" 2>&1 | tail -n 15
python - <<'PY'
import torch, time
for nbytes in (1<<20, 2<<20, 3<<20, 16<<20):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): d.copy_(h, non_blocking=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(20): d.copy_(h, non_blocking=True)
        e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("H2D pinned %d MiB: %.1f us  %.1f GB/s" % (nbytes >> 20, ms * 1e3, nbytes / ms / 1e6))
PY
