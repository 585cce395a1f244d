#!/bin/bash
# round 2, GPU call 25 (1 GPU): MLP backward v2 (64-column epilogues, one TMEM wait) — timing, full GPU suite, bench
mkdir -p gpurun_out
timeout 120 python scripts/time_mlp.py 1710000 > gpurun_out/r2_c25_time_mlp.txt 2>&1; echo "time_mlp rc=$?"; grep -v "Hash Enc" gpurun_out/r2_c25_time_mlp.txt | tail -4 | cut -c1-200
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c25_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_c25_pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c25_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_c25_smoke.log
timeout 600 python bench.py --steps 40 --warmup 5 --psnr-steps 0 --cpu-budget 1 > gpurun_out/r2_c25_bench_1gpu.json 2>/dev/null; echo "bench rc=$?"; grep '^{' gpurun_out/r2_c25_bench_1gpu.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['kernel_ms'], d['roofline']['samples'])"
