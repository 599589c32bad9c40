#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
ALIGNNET_B200_LIB=$PWD/tools/bin/libvar_tlap.so timeout 300 python tools/prof_step.py --workload c3 --steps 1 2>&1 | grep "^L2" > gpurun_out/r2_timeline_ap.txt; sed -n 9,16p gpurun_out/r2_timeline_ap.txt
bash tools/gpu_s21.sh ap32 ap100
