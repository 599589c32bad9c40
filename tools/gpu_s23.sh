#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_conv_stack.py tests/test_gpu_bf16.py tests/test_gpu_determinism.py tests/test_zz_fullsize_oracle.py -q -m gpu -x 2>&1 | tail -8
echo "=== bench c3"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>&1 | tail -1 > gpurun_out/r2_s23_bench.json; python -c "
import json; d=json.load(open('gpurun_out/r2_s23_bench.json')); print(d['ms_per_step'], d['value'], d['kernel_ms_by_tag'], d['gpu_launches'], d['roofline']['frac'])"
echo "=== timeline"; ALIGNNET_B200_LIB=$PWD/tools/bin/libvar_tl.so timeout 300 python tools/prof_step.py --workload c3 --steps 1 2>&1 | grep "^DG3\|^L2" > gpurun_out/r2_timeline2.txt; wc -l gpurun_out/r2_timeline2.txt
} > gpurun_out/r2_s23.log 2>&1
tail -30 gpurun_out/r2_s23.log | cut -c1-400
