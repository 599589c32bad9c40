#!/bin/bash
# round-2 final single-GPU artefacts: full GPU suite, smoke(), default bench line, ncu launch lists
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== full gpu suite"; time (timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15)
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "=== default bench"; timeout 900 python bench.py 2> gpurun_out/r2_final_bench.err | tail -1 > gpurun_out/r2_final_bench_c3.json; cut -c1-1500 gpurun_out/r2_final_bench_c3.json
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2_final_bench_reference.json; cut -c1-600 gpurun_out/r2_final_bench_reference.json
echo "=== ncu launch list c3"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_final_launches_c3.csv python tools/prof_step.py --workload c3 --steps 2 2>&1 | tail -1
echo "=== ncu launch list c2"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_final_launches_c2.csv python tools/prof_step.py --workload c2 --steps 3 2>&1 | tail -1
} > gpurun_out/r2_final1.log 2>&1
tail -60 gpurun_out/r2_final1.log | cut -c1-1600
