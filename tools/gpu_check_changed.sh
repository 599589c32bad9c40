#!/bin/bash
# One short call: from-clean compile on the GPU box, then the GPU tests of the code this session touched
# (optimiser dispatch, checkpoint slots, driver modes, provider writer) with their durations.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== from-clean build on the GPU box"
rm -rf alignnet-3d_b200/csrc/libalignnet_b200.so alignnet-3d_b200/csrc/build
( time python -c "import __graft_entry__ as g; g.build(); print('built')" ) 2>&1 | tail -5
ls -la alignnet-3d_b200/csrc/libalignnet_b200.so
echo "=== new tests"
timeout 600 python -m pytest tests/test_optimizer_momentum.py tests/test_provider.py tests/test_icp.py tests/test_tf_checkpoint.py -m gpu -q --durations=8 2>&1 | tail -25
timeout 600 python -m pytest tests/test_train_driver.py -m gpu -q --durations=8 2>&1 | tail -25
echo "=== engine paths touched"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bf16.py tests/test_abi.py -m gpu -q --durations=5 2>&1 | tail -15
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
} > gpurun_out/r2_check_changed.log 2>&1
cat gpurun_out/r2_check_changed.log | cut -c1-400
