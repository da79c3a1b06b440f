#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_inbatch.py tests/test_gpu_models.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/seg_tests.log
python - <<'PY' 2>&1 | tee gpurun_out/seg_bench.txt
import json, bench
r = bench.inbatch_trainer_steps()
for k, v in r.items():
    print(k, "ms/step %.4f (eager %.4f)" % (v["ms_per_step"], v["ms_per_step_eager_launches"]))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/twotower_launches.csv python tools/prof_twotower.py --steps 6 > gpurun_out/twotower.log 2>&1
python tools/prof_twotower.py --steps 6 --summarize gpurun_out/twotower_launches.csv | tee gpurun_out/twotower_summary.txt
