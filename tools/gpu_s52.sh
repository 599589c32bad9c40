#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python tools/debug_split_grads.py shipped_B32_N200 bf16x6 > gpurun_out/r2_s52_dbg.log 2>&1; echo rc=$?
head -70 gpurun_out/r2_s52_dbg.log
