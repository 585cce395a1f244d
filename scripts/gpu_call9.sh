#!/bin/bash
# round 2, GPU call 9 (2 GPUs): fp16 gradient transport (default) vs fp32; sharded optimizer (opt-in) correctness
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/check_dist_overlap.py > gpurun_out/r2_c9_dist_check.log 2>&1; echo "dist check rc=$?"; grep -v "Hash Enc" gpurun_out/r2_c9_dist_check.log | tail -4 | cut -c1-500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 30 --warmup 5 --psnr-steps 0 > gpurun_out/r2_c9_bench_2gpu_f16.json 2> gpurun_out/r2_c9_bench_2gpu_f16.err; echo "bench 2 f16 rc=$?"; tail -3 gpurun_out/r2_c9_bench_2gpu_f16.err
NGP_GRAD_F16=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 30 --warmup 5 --psnr-steps 0 > gpurun_out/r2_c9_bench_2gpu_f32.json 2> gpurun_out/r2_c9_bench_2gpu_f32.err; echo "bench 2 f32 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --psnr-steps 0 --cpu-budget 1 > gpurun_out/r2_c9_bench_1gpu.json 2>/dev/null; echo "bench 1 rc=$?"
for f in gpurun_out/r2_c9_bench_*.json; do echo "== $f"; grep '^{' $f | cut -c1-200; done
