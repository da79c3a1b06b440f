#!/bin/bash
# First multi-GPU call of round 2:   gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_first_call_2gpu.sh'
# libesr peer all-reduce / barrier kernels (fast_sync) and, on top of them, CUDA-graph capture of the sharded step.
mkdir -p gpurun_out
N=${N:-2}
ESR_TEST_EXPERIMENTAL=1 timeout 700 python -m pytest tests/test_gpu_sharded.py -q -k "fast_sync or overlap" > gpurun_out/r2_fast_sync_tests.log 2>&1
run() {  # name, extra flags
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $N --steps 100 --warmup 10 --no-cpu --no-inbatch --no-uniform $2 > gpurun_out/r2_bench_${N}gpu_$1.json 2> gpurun_out/r2_bench_${N}gpu_$1.err
}
run default ""
run fast_sync "--fast-sync"
run fast_sync_overlap "--fast-sync --overlap-ids"
run fast_sync_graphs "--fast-sync --step-graphs"
run fast_sync_overlap_graphs "--fast-sync --overlap-ids --step-graphs"
tail -3 gpurun_out/r2_fast_sync_tests.log
for f in gpurun_out/r2_bench_${N}gpu_*.json; do echo $f; head -c 400 $f; echo; done
# BASELINE configs[4]: 100M-row x 128 table row-sharded over the N GPUs -- lookup GB/s vs NVLink peak, full step, full-size properties
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 \
    tools/table_sweep.py --steps 30 --warmup 5 > gpurun_out/r2_table_sweep_${N}gpu.json 2> gpurun_out/r2_table_sweep_${N}gpu.err
head -c 1200 gpurun_out/r2_table_sweep_${N}gpu.json; tail -3 gpurun_out/r2_table_sweep_${N}gpu.err
python tools/r2_digest.py gpurun_out
