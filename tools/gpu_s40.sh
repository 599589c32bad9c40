#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
ALIGNNET_B200_LIB=$PWD/tools/bin/libvar_tlstg.so timeout 120 python tools/prof_step.py --workload c3 --steps 1 2>&1 | grep "^L2" > gpurun_out/r2_timeline_stg.txt; sed -n 9,16p gpurun_out/r2_timeline_stg.txt
bash tools/gpu_s31.sh stg
