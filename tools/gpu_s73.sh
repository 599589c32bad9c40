#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_s73_bench_2gpu.json 2> gpurun_out/r2_s73_bench.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_s73_bench_2gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['collective'])
for k,v in d.get('configs',{}).items(): print(k, v['n_gpus'], round(v['ms_per_step'],3), round(v['value']), round(v['e2e']['value']))
PY
tail -3 gpurun_out/r2_s73_bench.err
timeout 600 python -m pytest tests/test_gpu_dp.py -q -m gpu -x 2>&1 | tail -3
