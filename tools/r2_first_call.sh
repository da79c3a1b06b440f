#!/bin/bash
# First gpurun call of round 2 (1 GPU): everything written at the end of round 1 without GPU time.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/r2_first_call.sh'
# Results land in gpurun_out/r2_*.  Nothing here is a benchmark of record except the default bench.py line.
mkdir -p gpurun_out
export ESR_TEST_EXPERIMENTAL=1
# 1. parity of the experimental pieces (bit-identity of the accreg row pass, the reference-run pipeline through libesr)
timeout 600 python -m pytest tests -q -m gpu --deselect tests/test_gpu_virtual_peers.py -rs > gpurun_out/r2_accreg_tests.log 2>&1
cp gpurun_out/r2_accreg_tests.log gpurun_out/r2_ref_golden.log
# 1b. N virtual ranks on this one GPU: the peer path's integer kernels vs oracle/index.py, the sharded step vs the oracle
timeout 400 python -m pytest tests/test_gpu_virtual_peers.py -q > gpurun_out/r2_virtual_peers.log 2>&1
# 2. row-pass A/B, default vs accreg (Zipf + uniform, checksums must match)
timeout 150 python tools/probe_l2_hints.py --variants 0,3,4,5 --out gpurun_out/r2_probe_accreg.json > gpurun_out/r2_probe_accreg.log 2>&1
# 3. bench lines, default and accreg (no CPU leg, no in-batch leg: short)
unset ESR_TEST_EXPERIMENTAL
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --no-inbatch > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --no-inbatch --kernel accreg > gpurun_out/r2_bench_accreg.json 2> gpurun_out/r2_bench_accreg.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --no-inbatch --kernel hot > gpurun_out/r2_bench_hot.json 2> gpurun_out/r2_bench_hot.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --no-inbatch --no-uniform --stream-priority > gpurun_out/r2_bench_priority.json 2> gpurun_out/r2_bench_priority.err
# 4. DRAM traffic of the row pass on the uniform stream, default vs accreg (one ncu pass each, row-pass kernel only)
for k in auto accreg; do
  timeout 300 ncu --set full --clock-control none -k regex:k_glove_rows_grp_async -c 2 --csv --page raw \
      --log-file gpurun_out/r2_ncu_rows_${k}.csv python tools/prof_glove.py --V 1000000 --D 128 --B 262144 --uniform --steps 1 \
      --variant $([ $k = accreg ] && echo 3 || echo 0) > gpurun_out/r2_ncu_rows_${k}.log 2>&1
done
# 5. the Zipf stream (the headline): default vs hot-row cache, one ncu pass each (L1/shared pipe, L2 reads, duration)
for v in 0 4; do
  timeout 300 ncu --set full --clock-control none -k regex:k_glove_rows_grp_async -c 2 --csv --page raw \
      --log-file gpurun_out/r2_ncu_rows_zipf_v${v}.csv python tools/prof_glove.py --V 1000000 --D 128 --B 262144 --steps 1 \
      --variant $v > gpurun_out/r2_ncu_rows_zipf_v${v}.log 2>&1
done
tail -3 gpurun_out/r2_accreg_tests.log gpurun_out/r2_ref_golden.log gpurun_out/r2_virtual_peers.log
cat gpurun_out/r2_probe_accreg.json | head -c 1500
python tools/r2_digest.py gpurun_out
