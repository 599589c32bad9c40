#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== full gpu suite"; time (timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40)
} > gpurun_out/r2_s11.log 2>&1
tail -60 gpurun_out/r2_s11.log | cut -c1-500
