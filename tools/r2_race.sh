#!/bin/bash
mkdir -p gpurun_out
timeout 240 compute-sanitizer --tool racecheck python tools/sanitize_new_kernels.py 2>&1 | tail -6 | tee gpurun_out/r2_final_racecheck.txt
