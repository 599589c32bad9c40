#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
echo "=== full gpu suite"
( time timeout 1500 python -m pytest tests/ -x -q -m gpu -s ) > gpurun_out/r2_s62_pytest.log 2>&1; echo "rc=$?"
grep -v "^$" gpurun_out/r2_s62_pytest.log | tail -12 | cut -c1-300
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -3
