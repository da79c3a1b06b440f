#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_sharded_topk.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -6 | tee gpurun_out/sharded_topk_2gpu.txt
