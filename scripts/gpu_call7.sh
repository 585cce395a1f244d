#!/bin/bash
# round 2, GPU call 7 (1 GPU): suite with the inf flag at the source, MLP backward error measurement, frame800 (leap with
# per-lane binade validity), lego bench, grouped hash backward cost at one rank
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_c7_gputests.log 2>&1; echo "suite rc=$?"; tail -6 gpurun_out/r2_c7_gputests.log
timeout 300 python scripts/measure_mlp_bwd_error.py > gpurun_out/r2_mlp_bwd_error.txt 2>&1; cat gpurun_out/r2_mlp_bwd_error.txt | grep -v "Hash Encoder"
timeout 900 python bench.py --steps 20 --warmup 5 --config frame800 > gpurun_out/r2_c7_bench_frame800.json 2> gpurun_out/r2_c7_bench_frame800.err; echo "frame rc=$?"
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_c7_bench_lego_half.json 2> gpurun_out/r2_c7_bench_lego_half.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_c7_bench_lego_half.err
NGP_AR_OVERLAP=1 NGP_AR_FORCE=1 timeout 900 python bench.py --steps 30 --warmup 5 --psnr-steps 0 --cpu-budget 1 > gpurun_out/r2_c7_bench_lego_half_grouped.json 2>/dev/null; echo "bench grouped rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --profile-from-start off \
    --csv --log-file gpurun_out/r2_launches_frame800.csv python bench.py --config frame800 --ncu-window 1 > gpurun_out/r2_c7_ncu_frame.log 2>&1; echo "ncu frame launch list rc=$?"
for f in gpurun_out/r2_c7_bench_*.json; do echo "== $f"; grep '^{' $f | cut -c1-200; done
