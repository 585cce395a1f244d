#!/bin/bash
# round 2, GPU call 13 (2 GPUs): peer-memory optimizer step (csrc/p2p.cu) — correctness vs the NCCL variants, 2-GPU bench A/B
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/check_dist_overlap.py > gpurun_out/r2_c13_dist_check.log 2>&1; echo "dist check rc=$?"; grep -v "Hash Enc" gpurun_out/r2_c13_dist_check.log | tail -6 | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 40 --warmup 5 --psnr-steps 0 > gpurun_out/r2_c13_bench_2gpu_p2p.json 2> gpurun_out/r2_c13_bench_2gpu_p2p.err; echo "bench 2 p2p rc=$?"; grep -v "Hash Enc" gpurun_out/r2_c13_bench_2gpu_p2p.err | tail -3
NGP_P2P_ADAM=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 40 --warmup 5 --psnr-steps 0 > gpurun_out/r2_c13_bench_2gpu_nccl.json 2> gpurun_out/r2_c13_bench_2gpu_nccl.err; echo "bench 2 nccl rc=$?"
timeout 600 python bench.py --gpus 1 --steps 40 --warmup 5 --psnr-steps 0 --cpu-budget 1 > gpurun_out/r2_c13_bench_1gpu.json 2>/dev/null; echo "bench 1 rc=$?"
for f in gpurun_out/r2_c13_bench_*.json; do echo "== $f"; grep '^{' $f | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['parallelism'][:60], d['config']['samples_per_ray'])"; done
