#!/bin/bash
# round 2, GPU call 2: full GPU suite on the new trainer / grid update / tests, bench in every configuration
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_c2_smi.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2_c2_gputests.log 2>&1; echo "suite rc=$?"
tail -15 gpurun_out/r2_c2_gputests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_c2_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_c2_bench_lego_half.json 2> gpurun_out/r2_c2_bench_lego_half.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_c2_bench_lego_half.err
timeout 900 python bench.py --steps 20 --warmup 5 --config frame800 > gpurun_out/r2_c2_bench_frame800.json 2> gpurun_out/r2_c2_bench_frame800.err; echo "frame rc=$?"; tail -3 gpurun_out/r2_c2_bench_frame800.err
timeout 900 python bench.py --steps 20 --warmup 5 --config garden16 --psnr-steps 0 > gpurun_out/r2_c2_bench_garden16.json 2> gpurun_out/r2_c2_bench_garden16.err; echo "garden rc=$?"; tail -3 gpurun_out/r2_c2_bench_garden16.err
timeout 600 python bench.py --steps 20 --warmup 5 --config lego_fp32_1024 --psnr-steps 0 > gpurun_out/r2_c2_bench_lego_fp32_1024.json 2> gpurun_out/r2_c2_bench_lego_fp32_1024.err; echo "fp32 rc=$?"; tail -3 gpurun_out/r2_c2_bench_lego_fp32_1024.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_c2_ref_lego_half.json 2> gpurun_out/r2_c2_ref_lego_half.err; echo "ref rc=$?"
for f in gpurun_out/r2_c2_bench_*.json gpurun_out/r2_c2_ref_*.json; do echo "== $f"; cut -c1-400 $f; done
