#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_glove.py tests/test_gpu_virtual_peers.py -x -q -m gpu -k "plan or routed" 2>&1 | tail -3 | tee gpurun_out/sort_tests.log
timeout 120 python tools/prof_plan.py 2>&1 | tail -3 | tee gpurun_out/sort_prof.json
timeout 120 python tools/prof_plan.py --trace > gpurun_out/sort_trace.txt 2>&1
grep -A10 "^heads" gpurun_out/sort_trace.txt | grep -v "start time"
