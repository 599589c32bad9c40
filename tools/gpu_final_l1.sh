#!/bin/bash
# Final validation of the tree + refreshed evidence: full GPU suite, smoke, default bench, launch list, ncu --set full of the
# forward kernels and the statistics pass.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
echo "=== full gpu suite"
( time timeout 900 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r2d_pytest.log 2>&1; echo "rc=$?"
tail -4 gpurun_out/r2d_pytest.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== default bench"
timeout 600 python bench.py > gpurun_out/r2d_bench_c3.json 2> gpurun_out/r2d_bench.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d_bench_c3.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'], d['kernel_ms_by_tag'])
for k,v in d.get('configs',{}).items(): print(k, round(v['ms_per_step'],3), round(v['value']), round(v['e2e']['value']))
PY
echo "=== ncu launch list c3"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches_c3.csv python tools/prof_step.py --workload c3 --steps 2 > /dev/null 2>&1; echo rc=$?
echo "=== ncu full fwd + stats2"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:conv_stack_fwd_kernel|conv_stats2_kernel" -c 12 -o gpurun_out/r2d_fwd -f python tools/prof_step.py --workload c3 --steps 1 2>&1 | tail -1
