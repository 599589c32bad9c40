#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_split.py > gpurun_out/r2_s74_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -c "Invalid\|out of bounds" gpurun_out/r2_s74_memcheck.log; tail -4 gpurun_out/r2_s74_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_split.py > gpurun_out/r2_s74_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/r2_s74_racecheck.log
