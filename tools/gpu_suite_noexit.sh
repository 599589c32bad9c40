#!/bin/bash
# The full GPU suite WITHOUT -x (every failure listed), then the convergence-equivalence test repeated (its margins).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -q -m gpu --durations=15 ) > gpurun_out/r2_suite_all.log 2>&1; echo "suite rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|^real" gpurun_out/r2_suite_all.log | cut -c1-300
grep -A16 "slowest 15" gpurun_out/r2_suite_all.log | cut -c1-200
for i in 1 2 3; do
  timeout 300 python -m pytest tests/test_gpu_convergence.py -q -m gpu -s -k bf16_training 2>&1 | grep -E "windows:|train loss:|per_checkpoint|passed|failed|^E " | cut -c1-420
done > gpurun_out/r2_convergence_repeats.log 2>&1
cat gpurun_out/r2_convergence_repeats.log
