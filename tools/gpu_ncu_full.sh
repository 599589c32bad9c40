#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== ncu full backward kernels S2 b0"; timeout 900 ncu --set full --clock-control none --import-source on -k "regex:dgrad3_kernel|bwd_l2_kernel" -s 4 -c 2 -o gpurun_out/r2g_bwd -f python tools/prof_step.py --workload c3 --steps 1 2>&1 | tail -2
echo "=== ncu full fwd kernels"; timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_stack_fwd_kernel" -c 6 -o gpurun_out/r2g_fwd -f python tools/prof_step.py --workload c3 --steps 1 2>&1 | tail -2
} > gpurun_out/r2_ncu_full.log 2>&1
tail -12 gpurun_out/r2_ncu_full.log | cut -c1-300
