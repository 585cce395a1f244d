#!/bin/bash
# round 2, GPU call 15 (8 GPUs): peer-memory optimizer step at N = 8 — correctness check, bench p2p vs NCCL fp16
mkdir -p gpurun_out
N=${NGPUS:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 scripts/check_dist_overlap.py > gpurun_out/r2_c15_dist_check_${N}.log 2>&1; echo "dist check rc=$?"; grep -v "Hash Enc" gpurun_out/r2_c15_dist_check_${N}.log | tail -3 | cut -c1-600
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 40 --warmup 5 --psnr-steps 0 > gpurun_out/r2_c15_bench_${N}gpu_p2p.json 2> gpurun_out/r2_c15_bench_${N}gpu_p2p.err; echo "bench p2p rc=$?"; grep -v "Hash Enc" gpurun_out/r2_c15_bench_${N}gpu_p2p.err | grep -i "error\|Traceback" | head -5
NGP_P2P_ADAM=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 40 --warmup 5 --psnr-steps 0 > gpurun_out/r2_c15_bench_${N}gpu_nccl.json 2> gpurun_out/r2_c15_bench_${N}gpu_nccl.err; echo "bench nccl rc=$?"
timeout 300 python bench.py --gpus 1 --steps 40 --warmup 5 --psnr-steps 0 --cpu-budget 1 > gpurun_out/r2_c15_bench_1gpu.json 2>/dev/null; echo "bench 1 rc=$?"
for f in gpurun_out/r2_c15_bench_*.json; do echo "== $f"; grep '^{' $f | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['parallelism'][:60], d['config']['samples_per_ray'])"; done
