#!/bin/bash
# round 2, GPU call 33 (2 GPUs): the 2-GPU bench line with the final code
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 40 --warmup 5 --psnr-steps 0 > gpurun_out/r2_bench_final_2gpu.json 2> gpurun_out/r2_c33_bench_2gpu.err; echo "bench 2 rc=$?"
grep '^{' gpurun_out/r2_bench_final_2gpu.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], d['config']['parallelism'][:50])"
