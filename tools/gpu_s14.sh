#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== driver test"; timeout 600 python -m pytest tests/test_train_driver.py -q -m gpu -x 2>&1 | grep -v "^$" | tail -40
echo "=== DP hardware test (2 GPUs)"; timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -s 2>&1 | tail -30
echo "=== bench 2 GPUs"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-configs 2>&1 | tail -1 > gpurun_out/r2_s14_bench2.json; python -c "
import json; d=json.load(open('gpurun_out/r2_s14_bench2.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['collective'])"
} > gpurun_out/r2_s14.log 2>&1
tail -100 gpurun_out/r2_s14.log | cut -c1-1200
