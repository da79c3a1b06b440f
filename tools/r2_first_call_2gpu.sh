#!/bin/bash
# First multi-GPU call of round 2:   gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_first_call_2gpu.sh'
mkdir -p gpurun_out
N=${N:-2}
# 0. NVLink row-transfer shapes (pull / push / bulk), one and both directions
timeout 120 tools/microbench/nvlink_rows > gpurun_out/r2_nvlink_rows.txt 2>&1
# 1. parity of the peer path with the libesr sync kernels / id overlap
ESR_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_sharded.py -q -x > gpurun_out/r2_fast_sync_tests.log 2>&1
run() {  # name, extra flags
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $N --steps 100 --warmup 10 --no-cpu --no-inbatch --no-uniform $2 > gpurun_out/r2_bench_${N}gpu_$1.json 2> gpurun_out/r2_bench_${N}gpu_$1.err
}
run default ""
run fast_sync "--fast-sync"
run fast_sync_overlap "--fast-sync --overlap-ids"
run fast_sync_overlap_graphs "--fast-sync --overlap-ids --step-graphs"
tail -n 3 gpurun_out/r2_fast_sync_tests.log
for f in gpurun_out/r2_bench_${N}gpu_*.json; do echo $f; head -c 400 $f; echo; done
# BASELINE configs[4]: 100M-row x 128 table row-sharded over the N GPUs
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 \
    tools/table_sweep.py --steps 20 --warmup 3 > gpurun_out/r2_table_sweep_${N}gpu.json 2> gpurun_out/r2_table_sweep_${N}gpu.err
head -c 1500 gpurun_out/r2_table_sweep_${N}gpu.json; tail -n 3 gpurun_out/r2_table_sweep_${N}gpu.err
cat gpurun_out/r2_nvlink_rows.txt
