#!/bin/bash
# round 2, GPU call 12 (1 GPU): kernel-level A/B (gather widths, persistent grids) + co-run efficiency of hash || MLP; timelines
mkdir -p gpurun_out
timeout 900 python scripts/time_hash.py > gpurun_out/r2_c12_time_hash.jsonl 2> gpurun_out/r2_c12_time_hash.err; echo "time_hash rc=$?"; tail -3 gpurun_out/r2_c12_time_hash.err; grep '^{' gpurun_out/r2_c12_time_hash.jsonl | cut -c1-260
timeout 300 python scripts/step_timeline.py gpurun_out/r2_c12_timeline_k1.txt > /dev/null 2> gpurun_out/r2_c12_tl1.err; echo "timeline k1 rc=$?"; head -1 gpurun_out/r2_c12_timeline_k1.txt
NGP_STEP_CHUNKS=4 timeout 300 python scripts/step_timeline.py gpurun_out/r2_c12_timeline_k4.txt > /dev/null 2> gpurun_out/r2_c12_tl4.err; echo "timeline k4 rc=$?"; head -1 gpurun_out/r2_c12_timeline_k4.txt
