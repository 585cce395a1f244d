#!/bin/bash
# round 2, GPU call 30 (2 GPUs): dist check with the final kernels
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/check_dist_overlap.py > gpurun_out/r2_c30_dist_check.log 2>&1; echo "dist check rc=$?"; grep -v "Hash Enc" gpurun_out/r2_c30_dist_check.log | grep "dist check ok\|AssertionError" | head -3 | cut -c1-400
