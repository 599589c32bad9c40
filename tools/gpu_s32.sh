#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== tests"; timeout 600 python -m pytest tests/test_gpu_conv_stack.py tests/test_gpu_bf16.py tests/test_gpu_determinism.py tests/test_zz_fullsize_oracle.py tests/test_gpu_eval_cache.py -q -m gpu -x 2>&1 | tail -4
ALIGNNET_B200_LIB=$PWD/tools/bin/libvar_ftl.so timeout 300 python tools/prof_step.py --workload c3 --steps 1 2>&1 | grep "^FWD" > gpurun_out/r2_timeline_fwd2.txt; wc -l gpurun_out/r2_timeline_fwd2.txt
} > gpurun_out/r2_s32.log 2>&1
cat gpurun_out/r2_s32.log | cut -c1-300
bash tools/gpu_s21.sh fwdprev
