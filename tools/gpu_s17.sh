#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== ncu launch list c3 (fc2 + pack)"; timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -k regex:'fc2_gemm|pack_kernel' --csv --log-file gpurun_out/r2d_fc_launches.csv python tools/prof_step.py --workload c3 --steps 2 2>&1 | tail -1
echo "=== ncu full: head fc1 fwd"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:fc2_gemm -s 6 -c 1 -o gpurun_out/r2d_fc2 -f python tools/prof_step.py --workload c3 --steps 1 2>&1 | tail -1
} > gpurun_out/r2_s17.log 2>&1
tail -5 gpurun_out/r2_s17.log
