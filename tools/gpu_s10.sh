#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== full gpu suite"; time (timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15)
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
} > gpurun_out/r2_s10.log 2>&1
tail -40 gpurun_out/r2_s10.log | cut -c1-600
