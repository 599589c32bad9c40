#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_s54_launches_x6.csv python tools/prof_step.py --workload c3 --steps 1 --precision bf16x6 > gpurun_out/r2_s54.log 2>&1; echo rc=$?
python tools/launch_summary.py gpurun_out/r2_s54_launches_x6.csv 1 | head -40
