#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_determinism.py tests/test_gpu_split_gemm.py tests/test_gpu_parity.py tests/test_zz_fullsize_oracle.py -q -m gpu > gpurun_out/r2_s65.log 2>&1; echo "rc=$?"
tail -4 gpurun_out/r2_s65.log | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_s65_bench.json 2> gpurun_out/r2_s65_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_s65_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for k,v in d.get('configs',{}).items(): print(k, v['ms_per_step'], v['value'])
PY
tail -3 gpurun_out/r2_s65_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_s65_launches_x6.csv python tools/prof_step.py --workload c3 --steps 1 --precision bf16x6 > gpurun_out/r2_s65.nlog 2>&1; echo rc=$?
python tools/launch_summary.py gpurun_out/r2_s65_launches_x6.csv 1 2>/dev/null | head -12
