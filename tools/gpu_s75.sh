#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool initcheck --print-limit 40 python tools/sanitize_split.py > gpurun_out/r2_s75_initcheck.log 2>&1; echo "initcheck rc=$?"
grep "Uninitialized\|at .*an3d\|ERROR SUMMARY" gpurun_out/r2_s75_initcheck.log | sort | uniq -c | sort -rn | head -30
