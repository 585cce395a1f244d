#!/bin/bash
# round 2, GPU call 29 (2 GPUs): final multi-GPU sanity with the final kernels — dist check, 2-GPU bench, 1-GPU bench, frame800
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/check_dist_overlap.py > gpurun_out/r2_c29_dist_check.log 2>&1; echo "dist check rc=$?"; grep -v "Hash Enc" gpurun_out/r2_c29_dist_check.log | tail -1 | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 40 --warmup 5 --psnr-steps 0 > gpurun_out/r2_bench_final_2gpu.json 2> gpurun_out/r2_c29_bench_2gpu.err; echo "bench 2 rc=$?"
timeout 300 python bench.py --gpus 1 --steps 40 --warmup 5 --psnr-steps 0 --cpu-budget 1 > gpurun_out/r2_c29_bench_1gpu.json 2>/dev/null; echo "bench 1 rc=$?"
timeout 300 python bench.py --config frame800 --steps 30 --warmup 5 --cpu-budget 2 > gpurun_out/r2_bench_final_frame800.json 2>/dev/null; echo "frame800 rc=$?"
for f in gpurun_out/r2_bench_final_2gpu.json gpurun_out/r2_c29_bench_1gpu.json gpurun_out/r2_bench_final_frame800.json; do echo "== $f"; grep '^{' $f | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], (d.get('roofline') or {}).get('traffic'))"; done
