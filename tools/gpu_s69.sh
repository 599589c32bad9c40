#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_fullsize_oracle.py -q -m gpu > gpurun_out/r2_s69.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/r2_s69.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_s69_launches_x6.csv python tools/prof_step.py --workload c3 --steps 1 --precision bf16x6 > gpurun_out/r2_s69.nlog 2>&1; echo rc=$?
python tools/launch_summary.py gpurun_out/r2_s69_launches_x6.csv 1 2>/dev/null | grep "apply\|kernels,"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_s69_launches_cdef512.csv python tools/prof_step.py --workload cdef512 --steps 1 --precision bf16 > gpurun_out/r2_s69.nlog 2>&1; echo rc=$?
python tools/launch_summary.py gpurun_out/r2_s69_launches_cdef512.csv 1 2>/dev/null | grep "apply\|kernels,"
