#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== tests (quick set)"; timeout 900 python -m pytest tests/test_gpu_conv_stack.py tests/test_gpu_bf16.py tests/test_gpu_determinism.py tests/test_gpu_convergence.py tests/test_gpu_eval_cache.py -q -m gpu -x -s 2>&1 | grep -v "^$" | tail -30
echo "=== bench c3 (with configs)"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_s5_bench.json; cut -c1-3000 gpurun_out/r2_s5_bench.json
echo "=== ncu full EMB fwd"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_stack_fwd_kernel -s 4 -c 1 -o gpurun_out/r2b_fwd_emb -f python tools/prof_step.py --workload c3 --steps 1 2>&1 | tail -2
echo "=== ncu full stats2 EMB"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_stats2_kernel -s 4 -c 1 -o gpurun_out/r2b_stats2 -f python tools/prof_step.py --workload c3 --steps 1 2>&1 | tail -2
echo "=== ncu full dgrad3 + bwd_l2 + t1 (EMB = first of each in the backward)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dgrad3_kernel|bwd_l2_kernel|t1_sparse_kernel' -c 3 -o gpurun_out/r2b_bwd -f python tools/prof_step.py --workload c3 --steps 1 2>&1 | tail -2
echo "=== ncu launch list c3"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_c3.csv python tools/prof_step.py --workload c3 --steps 2 2>&1 | tail -1
} > gpurun_out/r2_s5.log 2>&1
tail -150 gpurun_out/r2_s5.log | cut -c1-600
