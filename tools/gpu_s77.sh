#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -q -s -m gpu -k layer_by_layer > gpurun_out/r2_s77.log 2>&1; echo "rc=$?"
grep "eval stable\|passed\|failed\|^E  " gpurun_out/r2_s77.log | cut -c1-300
