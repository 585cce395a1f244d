#!/bin/bash
# round 2, GPU call 1: v2 MLP forward parity + A/B timing, full GPU suite (validates the merged march fast path)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_c1_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_mlp_v2.py -q -m gpu > gpurun_out/r2_c1_v2test.log 2>&1; echo "v2test rc=$?"
timeout 300 python scripts/time_mlp.py 1710000 > gpurun_out/r2_c1_time_mlp.log 2>&1; echo "time_mlp rc=$?"
NGP_MLP_FWD=1 timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2_c1_gputests_v1.log 2>&1; echo "suite(v1 mlp) rc=$?"
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2_c1_gputests_auto.log 2>&1; echo "suite(auto) rc=$?"
NGP_MLP_FWD=1 timeout 600 python bench.py --steps 20 --warmup 5 --cpu-budget 4 > gpurun_out/r2_c1_bench_v1.log 2>&1; echo "bench v1 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-budget 4 > gpurun_out/r2_c1_bench_auto.log 2>&1; echo "bench auto rc=$?"
tail -3 gpurun_out/r2_c1_v2test.log; cat gpurun_out/r2_c1_time_mlp.log; tail -3 gpurun_out/r2_c1_gputests_v1.log; tail -3 gpurun_out/r2_c1_gputests_auto.log
