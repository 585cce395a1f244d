#!/bin/bash
# round 2, GPU call 3: compacting frame renderer (tests + fps A/B), ncu launch list + full capture of one graph step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2_c3_gputests.log 2>&1; echo "suite rc=$?"
tail -12 gpurun_out/r2_c3_gputests.log
timeout 900 python bench.py --steps 20 --warmup 5 --config frame800 > gpurun_out/r2_c3_bench_frame800.json 2> gpurun_out/r2_c3_bench_frame800.err; echo "frame rc=$?"; tail -3 gpurun_out/r2_c3_bench_frame800.err
NGP_FRAME_COMPACT=0 timeout 900 python bench.py --steps 20 --warmup 5 --config frame800 > gpurun_out/r2_c3_bench_frame800_nocompact.json 2> gpurun_out/r2_c3_bench_frame800_nocompact.err; echo "frame(no compaction) rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_c3_bench_lego_half.json 2> gpurun_out/r2_c3_bench_lego_half.err; echo "bench rc=$?"
# every launch of two steps with its device time (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --profile-from-start off \
    --csv --log-file gpurun_out/r2_launches_2steps.csv python bench.py --ncu-window 2 --psnr-steps 0 > gpurun_out/r2_c3_ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
# one step, full metric set, per graph node
timeout 1500 ncu --set full --clock-control none --import-source on --graph-profiling node --profile-from-start off \
    -o gpurun_out/r2_step_full python bench.py --ncu-window 1 --psnr-steps 0 > gpurun_out/r2_c3_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r2_step_full.ncu-rep --page raw --csv > gpurun_out/r2_step_full_raw.csv 2>/dev/null; echo "raw export rc=$?"
ls -la gpurun_out/*.ncu-rep
grep -h "ncu_window" gpurun_out/r2_c3_ncu_full.log gpurun_out/r2_c3_ncu_launch.log
for f in gpurun_out/r2_c3_bench_*.json; do echo "== $f"; cut -c1-300 $f; done
