#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
echo "=== tests (all but fullsize)"; timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_zz_fullsize_oracle.py --deselect tests/test_zz_fullsize_properties.py -x -s 2>&1 | grep -v "^$" | tail -25
echo "=== fullsize oracle"; timeout 900 python -m pytest tests/test_zz_fullsize_oracle.py -q -m gpu -s 2>&1 | tail -40
echo "=== fullsize properties"; timeout 600 python -m pytest tests/test_zz_fullsize_properties.py -q -m gpu 2>&1 | tail -15
echo "=== convergence probe"; timeout 600 python tools/convergence_probe.py --steps 300 2>&1 | tail -4
echo "=== bench c3"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-2500
echo "=== ncu full EMB fwd"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_stack_fwd_kernel -s 4 -c 1 -o gpurun_out/r2_fwd_emb -f python tools/prof_step.py --workload c3 --steps 1 2>&1 | tail -3
echo "=== ncu launch list c3"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_c3.csv python tools/prof_step.py --workload c3 --steps 2 2>&1 | tail -2
} > gpurun_out/r2_s4.log 2>&1
tail -150 gpurun_out/r2_s4.log | cut -c1-400
