#!/bin/bash
# round 2, GPU call 8 (2 GPUs): sharded optimizer - correctness + scaling vs the replicated update
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/check_dist_overlap.py > gpurun_out/r2_c8_dist_check.log 2>&1; echo "dist check rc=$?"; tail -4 gpurun_out/r2_c8_dist_check.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 30 --warmup 5 --psnr-steps 0 > gpurun_out/r2_c8_bench_2gpu_sharded.json 2> gpurun_out/r2_c8_bench_2gpu_sharded.err; echo "bench 2 sharded rc=$?"; tail -3 gpurun_out/r2_c8_bench_2gpu_sharded.err
NGP_SHARDED_ADAM=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 30 --warmup 5 --psnr-steps 0 > gpurun_out/r2_c8_bench_2gpu_replicated.json 2> gpurun_out/r2_c8_bench_2gpu_replicated.err; echo "bench 2 replicated rc=$?"
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --psnr-steps 0 --cpu-budget 1 > gpurun_out/r2_c8_bench_1gpu.json 2>/dev/null; echo "bench 1 rc=$?"
for f in gpurun_out/r2_c8_bench_*.json; do echo "== $f"; grep '^{' $f | cut -c1-200; done
