#!/bin/bash
# round 2, GPU call 18 (1 GPU): clock trace of the MLP backward v2
mkdir -p gpurun_out
timeout 300 python scripts/mlp_bwd_trace.py > gpurun_out/r2_c18_bwd_trace.txt 2>&1; echo "trace rc=$?"; cat gpurun_out/r2_c18_bwd_trace.txt | cut -c1-330
