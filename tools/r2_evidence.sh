#!/bin/bash
# Round-2 evidence run (1 GPU): ncu launch lists + full captures of the dominant kernels, compute-sanitizer on small parity cases.
mkdir -p gpurun_out
# 1. launch list of the default bench command (shares of the step): every launch with its device time
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r2_launches_bench_b256k.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu --no-inbatch --no-table-100m --no-uniform > gpurun_out/r2_launches_bench.log 2>&1
# 2. full capture of the row pass, Zipf + uniform (DRAM traffic for roofline.traffic)
for u in "" "--uniform"; do
  timeout 300 ncu --set full --clock-control none -k regex:k_glove_rows_grp_async -s 2 -c 1 --csv --page raw \
      --log-file gpurun_out/r2_ncu_rows${u/--/_}.csv python tools/prof_glove.py --V 1000000 --D 128 --B 262144 $u --steps 2 > /dev/null 2>&1
done
# 3. sharded step at world = 1 (all "peers" local): launch list + full capture of the owner merge and the pair routing
timeout 400 ncu --target-processes all --metrics gpu__time_duration.sum --clock-control none -s 600 -c 300 --csv --log-file gpurun_out/r2_launches_routed_w1.csv \
    python -m torch.distributed.run --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29650 tools/prof_routed.py > gpurun_out/r2_launches_routed.log 2>&1
timeout 400 ncu --target-processes all --set full --clock-control none -k regex:k_peer_merge_adagrad -s 3 -c 1 --csv --page raw --log-file gpurun_out/r2_ncu_merge_w1.csv \
    python -m torch.distributed.run --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29651 tools/prof_routed.py > /dev/null 2>&1
# 4. fused top-k scan
timeout 200 ncu --set full --clock-control none -k regex:k_topk_scan -s 2 -c 1 --csv --page raw --log-file gpurun_out/r2_ncu_topk.csv python tools/prof_topk.py > /dev/null 2>&1
# 5. compute-sanitizer on small parity cases: the last-block-arrives reductions, the persistent row pass, the peer kernels
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest -q -x tests/test_gpu_glove.py tests/test_gpu_virtual_peers.py \
      -k "(adagrad_steps and 300-128-1024 and auto) or (pair_routing_kernels and 2) or (owner_routed_step and 3-300-128-512-reference)" \
      > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$?"; tail -n 4 gpurun_out/r2_sanitizer_$tool.log
done
ls -la gpurun_out | grep r2_ | tail -n 20
