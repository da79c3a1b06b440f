#!/bin/bash
mkdir -p gpurun_out
one() { # label, env, flags
  env $2 timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu --no-inbatch --no-table-100m --no-uniform $3 > gpurun_out/r2c15_$1.json 2> gpurun_out/r2c15_$1.err
  python - gpurun_out/r2c15_$1.json $1 <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-22s value %.3f G  ms %.4f  e2e %.3f G (%.4f ms)  rows %.1f us" % (sys.argv[2], d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"]*1e3))
except Exception as e:
    print(sys.argv[2], "parse error", e); print(open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
}
one rb_default "X=1" ""
one rb_full "X=1" "--row-blocks 0"
one rb_240 "X=1" "--row-blocks 240"
one rb_222 "X=1" "--row-blocks 222"
one rb_197 "X=1" "--row-blocks 197"
one rb_148 "X=1" "--row-blocks 148"
one sidehi_default "ESR_PIPE_SIDE_HI=1" ""
one sidehi_full "ESR_PIPE_SIDE_HI=1" "--row-blocks 0"
one sidehi_222 "ESR_PIPE_SIDE_HI=1" "--row-blocks 222"
