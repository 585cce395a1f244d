#!/bin/bash
# round 2, GPU call 14 (1 GPU): full GPU suite incl. the single-device tests of the peer-memory optimizer kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c14_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_c14_pytest.log
