#!/bin/bash
mkdir -p gpurun_out
for k in tma ldg fifo; do
  timeout 200 python bench.py --steps 60 --warmup 10 --no-cpu --no-inbatch --no-table-100m --kernel $k > gpurun_out/r2c19_bench_$k.json 2> gpurun_out/r2c19_bench_$k.err
  python - gpurun_out/r2c19_bench_$k.json $k <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-6s value %.3f G  ms %.4f  rows zipf %.1f us (frac %.3f)  uniform %.1f us (frac %.3f)" % (sys.argv[2], d["value"]/1e9, d["ms_per_step"], d["roofline"]["kernel_ms"]*1e3, d["roofline"]["frac"], d["roofline_uniform"]["kernel_ms"]*1e3, d["roofline_uniform"]["frac"]))
except Exception as e:
    print(sys.argv[2], "parse error", e); print(open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
timeout 300 ncu --set full --clock-control none -k regex:k_glove_rows_tma -s 2 -c 1 --csv --page raw --log-file gpurun_out/r2_ncu_rows_tma.csv python tools/prof_glove.py --V 1000000 --D 128 --B 262144 --impl 2 --steps 2 > /dev/null 2>&1
ls -la gpurun_out/r2_ncu_rows_tma.csv
