#!/bin/bash
# round-2 session 1: measure the three switches round 1 left "prepared, not measured"
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
for sw in AN3D_FWD_RING=1 AN3D_FWD_RING=2 AN3D_FWD_RING=3; do
  echo "=== $sw"; timeout 400 python tools/ab_env.py $sw --tests tests/test_gpu_conv_stack.py 2>&1 | tail -15
done
echo "=== AN3D_TWO_STREAMS=1"; timeout 300 python tools/ab_env.py AN3D_TWO_STREAMS=1 --tests 2>&1 | tail -12
echo "=== AN3D_TWO_STREAMS=1 c2"; timeout 300 python tools/ab_env.py AN3D_TWO_STREAMS=1 --workload c2 --tests 2>&1 | tail -12
echo "=== AN3D_EVAL_CACHE=1 c2"; AN3D_RUN_EXPERIMENTAL=1 timeout 300 python tools/ab_env.py AN3D_EVAL_CACHE=1 --workload c2 --tests tests/test_gpu_experimental.py 2>&1 | tail -12
echo "=== fp32 c3"; timeout 300 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -2
} > gpurun_out/r2_s1.log 2>&1
tail -120 gpurun_out/r2_s1.log
