#!/bin/bash
# round 2, GPU call 32 (1 GPU): the default bench line (all arms, PSNR leg, cpu_baseline) with the final code
mkdir -p gpurun_out
timeout 500 python bench.py > gpurun_out/r2_bench_final_1gpu.json 2> gpurun_out/r2_c32_bench.err; echo "bench rc=$?"
grep '^{' gpurun_out/r2_bench_final_1gpu.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d.get(k) for k in ('metric','value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'mod', d['e2e_modules']['value'], 'psnr', (d.get('psnr') or {}).get('psnr'), d['roofline']['kernel_ms'], d['roofline']['traffic'], d['clocks'])"
