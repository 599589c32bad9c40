#!/bin/bash
# 2-GPU call: the data-parallel hardware test (tests/dp_worker.py under torchrun) and the 2-GPU bench line.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== DP hardware test (2 GPUs)"; timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -s 2>&1 | grep -E "^\{|passed|failed|Error|assert|FAILED" | cut -c1-2500
echo "=== bench 2 GPUs"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-configs --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2c_bench_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_2gpu.json')); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['collective'])"
} > gpurun_out/r2c_dp2.log 2>&1
cat gpurun_out/r2c_dp2.log | cut -c1-2600
