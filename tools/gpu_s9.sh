#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== tmem microbench"; timeout 120 tools/bin/tmem_microbench 2>&1 | grep "pipelined"
echo "=== DP hardware test (2 GPUs)"; timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -s 2>&1 | tail -25
echo "=== bench 2 GPUs"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -3 > gpurun_out/r2_s9_bench2.json; cut -c1-3000 gpurun_out/r2_s9_bench2.json
echo "=== bench 1 GPU (same box)"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>&1 | tail -1 | cut -c1-300
} > gpurun_out/r2_s9.log 2>&1
tail -150 gpurun_out/r2_s9.log | cut -c1-3000
