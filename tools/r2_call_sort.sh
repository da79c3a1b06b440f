#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_glove.py -x -q -m gpu -k "plan or bench_shape or trainer" 2>&1 | tail -3 | tee gpurun_out/sort_tests.log
