#!/bin/bash
# What the driver runs at round end, in one call: the full GPU suite, smoke(), the default bench line and the reference arm.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
echo "=== full gpu suite"
( time timeout 1200 python -m pytest tests/ -x -q -m gpu --durations=12 ) > gpurun_out/r2_final_pytest.log 2>&1; echo "rc=$?"
tail -22 gpurun_out/r2_final_pytest.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== default bench"
timeout 900 python bench.py > gpurun_out/r2_final_bench_c3.json 2> gpurun_out/r2_final_bench.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench_c3.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])
for k,v in d.get('configs',{}).items(): print(k, round(v['ms_per_step'],3), round(v['value']), round(v['e2e']['value']))
print(d['cpu_baseline'])
PY
echo "=== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2>> gpurun_out/r2_final_bench.err; echo "rc=$?"; cut -c1-300 gpurun_out/r2_final_bench_reference.json
