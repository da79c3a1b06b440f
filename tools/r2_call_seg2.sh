#!/bin/bash
mkdir -p gpurun_out
for s in default cub default cub; do
if [ $s = cub ]; then export ESR_PLAN_SORT=cub; else unset ESR_PLAN_SORT; fi
python - <<'PY' 2>&1 | tee -a gpurun_out/seg_bench2.txt
import os, bench
r = bench.inbatch_trainer_steps()
print("plan sort:", os.environ.get("ESR_PLAN_SORT", "default (wide for in-batch)"))
for k, v in r.items():
    print("  ", k, "ms/step %.4f (eager %.4f)" % (v["ms_per_step"], v["ms_per_step_eager_launches"]))
PY
done
