#!/bin/bash
# round 2, GPU call 11 (1 GPU): wide (32-byte) hash gathers + chunk-pipelined step sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c11_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_c11_pytest.log
timeout 900 python scripts/chunk_sweep.py > gpurun_out/r2_c11_sweep.jsonl 2> gpurun_out/r2_c11_sweep.err; echo "sweep rc=$?"; tail -5 gpurun_out/r2_c11_sweep.err; grep '^{' gpurun_out/r2_c11_sweep.jsonl | cut -c1-330
