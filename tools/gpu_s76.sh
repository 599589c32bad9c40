#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/r2_s76.log 2>&1; echo "rc=$?"
tail -6 gpurun_out/r2_s76.log | cut -c1-250
