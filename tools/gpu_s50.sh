#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_split_gemm.py -x -q -s -m gpu > gpurun_out/r2_s50_gemm.log 2>&1; echo "gemm rc=$?"
tail -5 gpurun_out/r2_s50_gemm.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/r2_s50_parity.log 2>&1; echo "parity rc=$?"
tail -15 gpurun_out/r2_s50_parity.log
