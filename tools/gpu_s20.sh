#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_conv_stack.py tests/test_gpu_bf16.py tests/test_gpu_determinism.py tests/test_zz_fullsize_oracle.py -q -m gpu -x 2>&1 | tail -5
echo "=== bench c3"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>&1 | tail -1 > gpurun_out/r2_s20_bench.json; python -c "
import json; d=json.load(open('gpurun_out/r2_s20_bench.json')); print(d['ms_per_step'], d['value'], d['kernel_ms_by_tag'], d['gpu_launches'], d['roofline']['frac'])"
echo "=== ncu full backward kernels (EMB b0, EMB b1, S2 b0)"; timeout 900 ncu --set full --clock-control none --import-source on -k "regex:dgrad3_kernel|bwd_l2_kernel" -c 6 -o gpurun_out/r2f_bwd -f python tools/prof_step.py --workload c3 --steps 1 2>&1 | tail -2
echo "=== ncu launch list c3"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_c3.csv python tools/prof_step.py --workload c3 --steps 2 2>&1 | tail -1
} > gpurun_out/r2_s20.log 2>&1
tail -40 gpurun_out/r2_s20.log | cut -c1-600
