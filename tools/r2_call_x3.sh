#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/twotower_launches.csv python tools/prof_twotower.py --steps 6 > gpurun_out/twotower.log 2>&1
python tools/prof_twotower.py --steps 6 --summarize gpurun_out/twotower_launches.csv | tee gpurun_out/twotower_summary.txt
ESR_PLAN_TRACE=1 timeout 120 python tools/prof_plan.py --trace > gpurun_out/sort_trace.txt 2> gpurun_out/sort_trace.err; echo "trace rc=$?"; tail -3 gpurun_out/sort_trace.err
grep -v "start time by tile\|phase [0-9]:" gpurun_out/sort_trace.txt | head -30
