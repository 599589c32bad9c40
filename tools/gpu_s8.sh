#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== tmem microbench (with concurrent MMA)"; timeout 120 tools/bin/tmem_microbench 2>&1 | grep -v "depth 2" 
} > gpurun_out/r2_s8.log 2>&1
tail -60 gpurun_out/r2_s8.log
