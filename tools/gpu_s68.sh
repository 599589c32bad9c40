#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_s68_launches_cdef512.csv python tools/prof_step.py --workload cdef512 --steps 1 --precision bf16 > gpurun_out/r2_s68.nlog 2>&1; echo rc=$?
python tools/launch_summary.py gpurun_out/r2_s68_launches_cdef512.csv 1 2>/dev/null | head -24
