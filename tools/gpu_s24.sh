#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_conv_stack.py tests/test_gpu_bf16.py tests/test_gpu_determinism.py tests/test_zz_fullsize_oracle.py -q -m gpu -x 2>&1 | tail -4
echo "=== timeline"; ALIGNNET_B200_LIB=$PWD/tools/bin/libvar_tl.so timeout 300 python tools/prof_step.py --workload c3 --steps 1 2>&1 | grep "^DG3\|^L2" > gpurun_out/r2_timeline3.txt; wc -l gpurun_out/r2_timeline3.txt
ALIGNNET_B200_LIB=$PWD/tools/bin/libvar_tlap.so timeout 300 python tools/prof_step.py --workload c3 --steps 1 2>&1 | grep "^DG3\|^L2" > gpurun_out/r2_timeline3ap.txt; wc -l gpurun_out/r2_timeline3ap.txt
} > gpurun_out/r2_s24.log 2>&1
cat gpurun_out/r2_s24.log | cut -c1-300
bash tools/gpu_s21.sh ap
