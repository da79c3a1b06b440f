#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/sort_launches.csv python tools/prof_plan.py --eager 4 > /dev/null 2>&1
grep -v "^==" gpurun_out/sort_launches.csv | awk -F'","' 'NR>1{print $5, $(NF)}' | sed 's/(.*)//' | tail -16
ncu --set full --clock-control none --import-source on -k regex:"k_sort_pass|k_heads|k_sort_hist" -s 6 -c 6 -o gpurun_out/sort_full python tools/prof_plan.py --eager 3 > /dev/null 2>&1
ls -la gpurun_out/sort_full.ncu-rep
