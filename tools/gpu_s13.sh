#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_conv_stack.py tests/test_gpu_bf16.py tests/test_gpu_determinism.py tests/test_train_driver.py tests/test_zz_fullsize_oracle.py -q -m gpu -x 2>&1 | tail -5
echo "=== bench c3"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>&1 | tail -1 > gpurun_out/r2_s13_bench.json; python -c "
import json; d=json.load(open('gpurun_out/r2_s13_bench.json')); print(d['ms_per_step'], d['value'], d['kernel_ms_by_tag'], d['gpu_launches'], d['roofline']['frac'])"
echo "=== bench c3 single stream"; AN3D_TWO_STREAMS=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
} > gpurun_out/r2_s13.log 2>&1
tail -30 gpurun_out/r2_s13.log | cut -c1-400
