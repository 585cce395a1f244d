#!/bin/bash
# round 2, GPU call 24 (1 GPU): MLP backward v2 — three 64 KB slots, two outstanding weight-gradient requests per slot
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mlp_bwd" > gpurun_out/r2_c24_pytest_mlp.log 2>&1; echo "pytest mlp_bwd rc=$?"; tail -3 gpurun_out/r2_c24_pytest_mlp.log | cut -c1-200
timeout 120 python scripts/time_mlp.py 1710000 > gpurun_out/r2_c24_time_mlp.txt 2>&1; echo "time_mlp rc=$?"; grep -v "Hash Enc" gpurun_out/r2_c24_time_mlp.txt | tail -5 | cut -c1-200
timeout 100 python scripts/mlp_bwd_trace.py > gpurun_out/r2_c24_bwd_trace.txt 2>&1; echo "trace rc=$?"; head -11 gpurun_out/r2_c24_bwd_trace.txt | cut -c1-330
