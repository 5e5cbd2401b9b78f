#!/bin/bash
# build a timing variant of the library: routing_fused2.cu compiled with extra -D flags, linked with the
# objects of the regular build (python -m qecnetworks_b200.build first) -> variants/libqec_NAME.so
#   scripts/build_variant.sh NAME -DFOO -DBAR=1 ; QEC_LIB=variants/libqec_NAME.so python scripts/time_pool2.py
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants/obj
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -ftz=true -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" \
  -c qecnetworks_b200/csrc/routing_fused2.cu -o variants/obj/$name.o 2> variants/obj/$name.log
objs=$(ls qecnetworks_b200/build/*.o | grep -v routing_fused2.o)
nvcc -shared -o variants/libqec_$name.so variants/obj/$name.o $objs -gencode arch=compute_100a,code=sm_100a
grep -A2 "routing_fused2_bwd_kernelILi8ELb1" variants/obj/$name.log | grep -E "spill"
