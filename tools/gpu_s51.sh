#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_split_gemm.py -q -m gpu > gpurun_out/r2_s51_gemm.log 2>&1; echo "gemm rc=$?"
tail -3 gpurun_out/r2_s51_gemm.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/r2_s51_parity.log 2>&1; echo "parity rc=$?"
tail -8 gpurun_out/r2_s51_parity.log
timeout 1200 python -m pytest tests/test_zz_fullsize_oracle.py -q -s -m gpu > gpurun_out/r2_s51_full.log 2>&1; echo "fullsize rc=$?"
tail -12 gpurun_out/r2_s51_full.log
