#!/bin/bash
# usage: gpu_bench_ngpu.sh N   -- the bench at N GPUs (with the `configs` sub-object), builder-run record for profiles/
cd "$GRAFT_REPO_ROOT"
N=$1
mkdir -p gpurun_out
{
echo "=== bench $N GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2_bench_${N}gpu.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${N}gpu.json')); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['collective'], d['config']['host_binding']); print({k:(v['ms_per_step'], v['value'], v['e2e']['value']) for k,v in d.get('configs',{}).items()})"
if [ "$N" = "4" ]; then echo "=== DP hardware test"; timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -s 2>&1 | grep -E "^\{|passed|failed|Error|assert" | cut -c1-1500; fi
} > gpurun_out/r2_bench_ngpu_$N.log 2>&1
tail -30 gpurun_out/r2_bench_ngpu_$N.log | cut -c1-1600
