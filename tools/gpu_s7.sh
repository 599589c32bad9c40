#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== tests (conv stack, bf16, determinism)"; timeout 900 python -m pytest tests/test_gpu_conv_stack.py tests/test_gpu_bf16.py tests/test_gpu_determinism.py -q -m gpu -x 2>&1 | tail -5
echo "=== bench c3"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>&1 | tail -1 > gpurun_out/r2_s7_bench.json; cut -c1-400 gpurun_out/r2_s7_bench.json; python -c "
import json; d=json.load(open('gpurun_out/r2_s7_bench.json')); print(d['ms_per_step'], d['kernel_ms_by_tag'], d['gpu_launches'])"
echo "=== convergence probe 1600 steps"; timeout 900 python tools/convergence_probe.py --steps 1600 --every 200 2>&1 | tail -40
} > gpurun_out/r2_s7.log 2>&1
tail -80 gpurun_out/r2_s7.log | cut -c1-400
