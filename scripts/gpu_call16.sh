#!/bin/bash
# round 2, GPU call 16 (1 GPU): MLP backward v2 (four tile slots per persistent CTA) — parity tests, A/B timing, suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mlp_bwd" > gpurun_out/r2_c16_pytest_mlp.log 2>&1; echo "pytest mlp_bwd rc=$?"; tail -6 gpurun_out/r2_c16_pytest_mlp.log
timeout 300 python scripts/time_mlp.py 1710000 > gpurun_out/r2_c16_time_mlp.txt 2>&1; echo "time_mlp rc=$?"; grep -v "Hash Enc" gpurun_out/r2_c16_time_mlp.txt | tail -9
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c16_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_c16_pytest.log
timeout 600 python bench.py --steps 40 --warmup 5 --psnr-steps 0 --cpu-budget 1 > gpurun_out/r2_c16_bench_1gpu.json 2>/dev/null; echo "bench rc=$?"; grep '^{' gpurun_out/r2_c16_bench_1gpu.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['kernel_ms'], d['roofline']['samples'])"
