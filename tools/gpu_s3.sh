#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== tests"; timeout 600 python -m pytest tests/test_gpu_umma.py tests/test_gpu_conv_stack.py tests/test_gpu_bf16.py tests/test_gpu_eval_cache.py -x -q -m gpu 2>&1 | tail -25
echo "=== bench c3"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -3 | cut -c1-2500
echo "=== bench c2"; timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -3 | cut -c1-2500
} > gpurun_out/r2_s3.log 2>&1
tail -60 gpurun_out/r2_s3.log
