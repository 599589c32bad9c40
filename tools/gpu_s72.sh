#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_convergence.py -q -s -m gpu -k bf16x6 > gpurun_out/r2_s72.log 2>&1; echo "rc=$?"
grep "loss trajectory\|passed\|failed\|Error" gpurun_out/r2_s72.log | cut -c1-400
