#!/bin/bash
# round 2, GPU call 17 (1 GPU): MLP backward v2 with the descriptor table — parity, A/B timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mlp_bwd" > gpurun_out/r2_c17_pytest_mlp.log 2>&1; echo "pytest mlp_bwd rc=$?"; tail -3 gpurun_out/r2_c17_pytest_mlp.log
timeout 300 python scripts/time_mlp.py 1710000 > gpurun_out/r2_c17_time_mlp.txt 2>&1; echo "time_mlp rc=$?"; grep -v "Hash Enc" gpurun_out/r2_c17_time_mlp.txt | tail -5
