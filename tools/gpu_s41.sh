#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
ALIGNNET_B200_LIB=$PWD/tools/bin/libvar_ftl.so timeout 300 python tools/prof_step.py --workload c3 --steps 1 2>&1 | grep "^FWDM\|^FWDK" > gpurun_out/r2_timeline_fwdm.txt; cat gpurun_out/r2_timeline_fwdm.txt
