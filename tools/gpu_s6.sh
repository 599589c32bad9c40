#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== tests (quick set)"; timeout 900 python -m pytest tests/test_gpu_conv_stack.py tests/test_gpu_bf16.py tests/test_gpu_determinism.py tests/test_gpu_convergence.py -q -m gpu -x -s 2>&1 | grep -v "^$\|^trial" | tail -30
echo "=== bench c3 (with configs)"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_s6_bench.json; cut -c1-3500 gpurun_out/r2_s6_bench.json
echo "=== ncu full EMB fwd"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_stack_fwd_kernel -s 4 -c 1 -o gpurun_out/r2c_fwd_emb -f python tools/prof_step.py --workload c3 --steps 1 2>&1 | tail -2
echo "=== ncu launch list c3"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_c3.csv python tools/prof_step.py --workload c3 --steps 2 2>&1 | tail -1
} > gpurun_out/r2_s6.log 2>&1
tail -150 gpurun_out/r2_s6.log | cut -c1-3600
