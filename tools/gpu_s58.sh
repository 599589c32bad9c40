#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_astat -s 2 -c 1 -f -o gpurun_out/r2_s58_astat python tools/prof_step.py --workload c3 --steps 1 --precision bf16x6 > gpurun_out/r2_s58a.log 2>&1; echo rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_apply_split -s 0 -c 1 -f -o gpurun_out/r2_s58_apply python tools/prof_step.py --workload c3 --steps 1 --precision bf16x6 > gpurun_out/r2_s58b.log 2>&1; echo rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 25 -c 2 -f -o gpurun_out/r2_s58_bwd python tools/prof_step.py --workload c3 --steps 1 --precision bf16x6 > gpurun_out/r2_s58c.log 2>&1; echo rc=$?
ls -la gpurun_out/r2_s58*
