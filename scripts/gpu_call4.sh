#!/bin/bash
# round 2, GPU call 4: suite (round march parity, grouped backward), frame800 with the leap, deterministic lego bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_c4_gputests.log 2>&1; echo "suite rc=$?"
tail -8 gpurun_out/r2_c4_gputests.log
timeout 900 python bench.py --steps 20 --warmup 5 --config frame800 > gpurun_out/r2_c4_bench_frame800.json 2> gpurun_out/r2_c4_bench_frame800.err; echo "frame rc=$?"; tail -3 gpurun_out/r2_c4_bench_frame800.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_c4_bench_lego_half.json 2> gpurun_out/r2_c4_bench_lego_half.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_c4_bench_lego_half.err
timeout 900 python bench.py --steps 20 --warmup 5 --psnr-steps 0 > gpurun_out/r2_c4_bench_lego_half_b.json 2>/dev/null; echo "bench(repeat) rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --profile-from-start off \
    --csv --log-file gpurun_out/r2_launches_frame800.csv python bench.py --config frame800 --ncu-window 1 > gpurun_out/r2_c4_ncu_frame.log 2>&1; echo "ncu frame launch list rc=$?"
for f in gpurun_out/r2_c4_bench_*.json; do echo "== $f"; cut -c1-260 $f; done
