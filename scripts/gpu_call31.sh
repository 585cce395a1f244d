#!/bin/bash
# round 2, GPU call 31 (1 GPU): MLP forward v2 with converged-warp issue + constant descriptors; v1-vs-v2 backward test; full suite; bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c31_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_c31_pytest.log | cut -c1-200
timeout 120 python scripts/time_mlp.py 1710000 > gpurun_out/r2_c31_time_mlp.txt 2>&1; echo "time_mlp rc=$?"; grep -v "Hash Enc" gpurun_out/r2_c31_time_mlp.txt | tail -9 | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c31_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_c31_smoke.log
timeout 400 python bench.py --steps 40 --warmup 5 --psnr-steps 0 --cpu-budget 1 > gpurun_out/r2_c31_bench_1gpu.json 2>/dev/null; echo "bench rc=$?"; grep '^{' gpurun_out/r2_c31_bench_1gpu.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['kernel_ms'], d['roofline']['samples'])"
