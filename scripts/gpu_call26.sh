#!/bin/bash
# round 2, GPU call 26 (1 GPU): ncu evidence for the final kernels — launch list of two graph steps + full capture of one
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --profile-from-start off \
    --csv --log-file gpurun_out/r2_launches_2steps.csv python bench.py --ncu-window 2 --psnr-steps 0 > gpurun_out/r2_c26_ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --graph-profiling node --profile-from-start off \
    -o gpurun_out/r2_step_full python bench.py --ncu-window 1 --psnr-steps 0 > gpurun_out/r2_c26_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r2_step_full.ncu-rep --page raw --csv > gpurun_out/r2_ncu_step_full_raw.csv 2>/dev/null; echo "raw export rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
grep -h "ncu_window" gpurun_out/r2_c26_ncu_full.log gpurun_out/r2_c26_ncu_launch.log
