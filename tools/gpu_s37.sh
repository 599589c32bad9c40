#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_conv_stack.py tests/test_gpu_bf16.py tests/test_gpu_determinism.py tests/test_zz_fullsize_oracle.py -q -m gpu -x 2>&1 | tail -4
} > gpurun_out/r2_s37.log 2>&1
cat gpurun_out/r2_s37.log | cut -c1-300
bash tools/gpu_s21.sh "$@"
