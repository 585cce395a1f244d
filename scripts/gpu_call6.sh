#!/bin/bash
# round 2, GPU call 6 (2 GPUs): per-slice all-reduce behind the backward kernels - correctness + scaling
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/check_dist_overlap.py > gpurun_out/r2_c6_dist_check.log 2>&1; echo "dist check rc=$?"; tail -5 gpurun_out/r2_c6_dist_check.log
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --psnr-steps 0 > gpurun_out/r2_c6_bench_1gpu.json 2> gpurun_out/r2_c6_bench_1gpu.err; echo "bench 1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 30 --warmup 5 --psnr-steps 0 > gpurun_out/r2_c6_bench_2gpu.json 2> gpurun_out/r2_c6_bench_2gpu.err; echo "bench 2 rc=$?"; tail -3 gpurun_out/r2_c6_bench_2gpu.err
NGP_AR_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 30 --warmup 5 --psnr-steps 0 > gpurun_out/r2_c6_bench_2gpu_monolithic.json 2> gpurun_out/r2_c6_bench_2gpu_monolithic.err; echo "bench 2 (monolithic all-reduce) rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 --ref-budget 40 > gpurun_out/r2_c6_ref_2gpu.json 2>/dev/null; echo "ref under torchrun rc=$?"
for f in gpurun_out/r2_c6_bench_*.json gpurun_out/r2_c6_ref_2gpu.json; do echo "== $f"; grep '^{' $f | cut -c1-330; done
