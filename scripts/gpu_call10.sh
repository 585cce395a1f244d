#!/bin/bash
# round 2, GPU call 10 (1 GPU): state check after re-entry — full GPU suite, smoke, default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c10_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_c10_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c10_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_c10_smoke.log
timeout 900 python bench.py > gpurun_out/r2_c10_bench_1gpu.json 2> gpurun_out/r2_c10_bench_1gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_c10_bench_1gpu.err; grep '^{' gpurun_out/r2_c10_bench_1gpu.json | cut -c1-1500
