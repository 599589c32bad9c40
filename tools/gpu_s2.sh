#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== tmem microbench"; timeout 120 tools/bin/tmem_microbench
echo "=== AN3D_TWO_STREAMS=1 c2"; AN3D_TWO_STREAMS=1 timeout 200 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -30 | cut -c1-1500
} > gpurun_out/r2_s2.log 2>&1
tail -150 gpurun_out/r2_s2.log
