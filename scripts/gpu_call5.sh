#!/bin/bash
# round 2, GPU call 5: TMEM read bandwidth micro-benchmark, real timeline of the graph step, frame800 with the adaptive leap
mkdir -p gpurun_out
./scripts/micro/tmem_bw > gpurun_out/r2_tmem_bw.txt 2>&1; echo "tmem_bw rc=$?"; cat gpurun_out/r2_tmem_bw.txt
timeout 600 python scripts/step_timeline.py gpurun_out/r2_step_timeline.txt > gpurun_out/r2_c5_timeline.log 2>&1; echo "timeline rc=$?"; tail -45 gpurun_out/r2_step_timeline.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "round_march or fused_grid" > gpurun_out/r2_c5_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_c5_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --config frame800 > gpurun_out/r2_c5_bench_frame800.json 2> gpurun_out/r2_c5_bench_frame800.err; echo "frame rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --profile-from-start off \
    --csv --log-file gpurun_out/r2_launches_frame800.csv python bench.py --config frame800 --ncu-window 1 > gpurun_out/r2_c5_ncu_frame.log 2>&1; echo "ncu frame launch list rc=$?"
grep -E "march|composite_round" gpurun_out/r2_launches_frame800.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | head -20
cut -c1-260 gpurun_out/r2_c5_bench_frame800.json
