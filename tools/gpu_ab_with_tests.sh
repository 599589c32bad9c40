#!/bin/bash
# Same-box A/B of variant libraries (tools/build_variants.sh): parity tests of each variant, then the bench base / variants twice.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
for v in "$@"; do
  export ALIGNNET_B200_LIB=$PWD/tools/bin/libvar_$v.so
  echo "=== tests $v"; timeout 400 python -m pytest tests/test_gpu_conv_stack.py tests/test_gpu_bf16.py tests/test_gpu_determinism.py tests/test_gpu_eval_cache.py -q -m gpu -x 2>&1 | tail -3
done
for rep in 1 2; do
for v in base "$@"; do
  if [ "$v" = base ]; then unset ALIGNNET_B200_LIB; else export ALIGNNET_B200_LIB=$PWD/tools/bin/libvar_$v.so; fi
  echo -n "$v: "; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), {k: round(v,3) for k,v in d['kernel_ms_by_tag'].items()})
except Exception as e: print('FAILED', e)"
done
done
} > gpurun_out/r2c_ab.log 2>&1
cat gpurun_out/r2c_ab.log | cut -c1-300
