#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_determinism.py tests/test_gpu_split_gemm.py tests/test_gpu_parity.py tests/test_train_driver.py -q -m gpu -s > gpurun_out/r2_s64.log 2>&1; echo "rc=$?"
grep -v "^$" gpurun_out/r2_s64.log | grep -v "err/unit" | tail -8 | cut -c1-300
grep "vs fp32 on the same checkpoint" gpurun_out/r2_s64.log
for i in 1 2 3; do timeout 600 python -m pytest tests/test_zz_fullsize_oracle.py -q -s -m gpu -k default > gpurun_out/r2_s64_d$i.log 2>&1; grep "^default_arch_bf16 \|passed\|failed" gpurun_out/r2_s64_d$i.log | cut -c1-60,200-330; done
