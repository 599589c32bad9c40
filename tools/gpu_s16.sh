#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_fc_gemm.py tests/test_gpu_bf16.py tests/test_gpu_determinism.py tests/test_gpu_eval_cache.py tests/test_zz_fullsize_oracle.py -q -m gpu -x 2>&1 | tail -25
echo "=== bench c3"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>&1 | tail -1 > gpurun_out/r2_s16_bench.json; python -c "
import json; d=json.load(open('gpurun_out/r2_s16_bench.json')); print(d['ms_per_step'], d['value'], d['kernel_ms_by_tag'], d['gpu_launches'], d['roofline']['frac'])"
echo "=== bench c2"; timeout 600 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['kernel_ms_by_tag'], d['gpu_launches'])"
} > gpurun_out/r2_s16.log 2>&1
tail -40 gpurun_out/r2_s16.log | cut -c1-600
