#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== timeline"; ALIGNNET_B200_LIB=$PWD/tools/bin/libvar_tl.so timeout 300 python tools/prof_step.py --workload c3 --steps 1 2>&1 | grep "^DG3\|^L2" > gpurun_out/r2_timeline.txt; wc -l gpurun_out/r2_timeline.txt
} > gpurun_out/r2_s22.log 2>&1
cat gpurun_out/r2_s22.log
bash tools/gpu_s21.sh lossold
