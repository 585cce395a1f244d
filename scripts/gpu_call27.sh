#!/bin/bash
# round 2, GPU call 27 (1 GPU): final validation — full GPU suite, smoke, the default bench line and the other BASELINE configurations
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c27_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_c27_pytest.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c27_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_c27_smoke.log
timeout 600 python bench.py > gpurun_out/r2_bench_final_1gpu.json 2> gpurun_out/r2_c27_bench.err; echo "bench rc=$?"
for c in lego_fp32_1024 garden16 frame800; do
  timeout 400 python bench.py --config $c --steps 30 --warmup 5 --psnr-steps 0 --cpu-budget 2 > gpurun_out/r2_bench_final_$c.json 2>/dev/null; echo "bench $c rc=$?"
done
for f in gpurun_out/r2_bench_final_*.json; do echo "== $f"; grep '^{' $f | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d.get(k) for k in ('metric','value','ms_per_step')}, 'e2e', d['e2e']['value'], 'psnr', (d.get('psnr') or {}).get('psnr'), 'traffic', d.get('roofline',{}).get('traffic'))"; done
