#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
bash tools/gpu_s21.sh "$@"
{
for v in "$@"; do
  export ALIGNNET_B200_LIB=$PWD/tools/bin/libvar_$v.so; echo "=== tests $v"; timeout 600 python -m pytest tests/test_gpu_conv_stack.py tests/test_gpu_bf16.py -q -m gpu -x 2>&1 | tail -2
done
} > gpurun_out/r2_s31.log 2>&1
cat gpurun_out/r2_s31.log | cut -c1-300
