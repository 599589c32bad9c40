#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== ncu launch list c3"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_launches_c3.csv python tools/prof_step.py --workload c3 --steps 2 2>&1 | tail -1
echo "=== ncu launch list c2"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_launches_c2.csv python tools/prof_step.py --workload c2 --steps 3 2>&1 | tail -1
} > gpurun_out/r2_s18.log 2>&1
tail -5 gpurun_out/r2_s18.log
