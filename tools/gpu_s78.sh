#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_stack.py tests/test_gpu_bf16.py tests/test_gpu_determinism.py -q -m gpu > gpurun_out/r2_s78.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/r2_s78.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_s78_launches_c3.csv python tools/prof_step.py --workload c3 --steps 2 > /dev/null 2>&1; echo rc=$?
python tools/launch_summary.py gpurun_out/r2_s78_launches_c3.csv 2 2>/dev/null | grep "kernels,\|bwd_l1_finish\|pool_bwd_prep\|gw3_stats\|stats2_from\|sum_parts\|col_reduce\|bn_bwd_apply_img\|fold_acc\|gq_partial\|moments"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3', d['value'], d['ms_per_step'])"
