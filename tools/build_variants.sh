#!/bin/bash
# Builds tools/bin/libvar_<name>.so: the library with one translation unit recompiled under extra -D flags
# (A/B experiments on one GPU box: ALIGNNET_B200_LIB=tools/bin/libvar_x.so python bench.py ...).
#   tools/build_variants.sh <unit.cu> <name> <flags...>
set -e
cd "$(dirname "$0")/.."
python alignnet-3d_b200/build.py > /dev/null
unit=$1; name=$2; shift 2
C=alignnet-3d_b200/csrc
mkdir -p tools/bin
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -c $C/$unit -o tools/bin/var_$name.o
objs=$(ls $C/build/*.o | grep -v "/${unit%.cu}.o")
nvcc -shared -o tools/bin/libvar_$name.so $objs tools/bin/var_$name.o -lcudart
echo tools/bin/libvar_$name.so
