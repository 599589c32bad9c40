#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bf16.py tests/test_gpu_determinism.py tests/test_zz_fullsize_oracle.py tests/test_zz_fullsize_properties.py -q -m gpu -x 2>&1 | tail -15
echo "=== bench c3"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>&1 | tail -1 > gpurun_out/r2_s19_bench.json; python -c "
import json; d=json.load(open('gpurun_out/r2_s19_bench.json')); print(d['ms_per_step'], d['value'], d['kernel_ms_by_tag'], d['gpu_launches'], d['roofline']['frac'])"
} > gpurun_out/r2_s19.log 2>&1
tail -30 gpurun_out/r2_s19.log | cut -c1-600
