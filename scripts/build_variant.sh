#!/bin/bash
# build a timing variant of the library: both builds of the packed cluster kernels (routing_fused2.cu and
# routing_fused2_wide.cu) compiled with extra -D flags, linked with the other objects of the regular build
# (python -m qecnetworks_b200.build first) -> variants/libqec_NAME.so (git-ignored, travels with gpurun)
#   scripts/build_variant.sh NAME -DFOO -DBAR=1 ; QEC_LIB=variants/libqec_NAME.so python scripts/time_pool2.py
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants/obj
flags="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -ftz=true -std=c++17 -Xcompiler -fPIC -Xptxas -v"
nvcc $flags "$@" -c qecnetworks_b200/csrc/routing_fused2.cu -o variants/obj/$name.o 2> variants/obj/$name.log &
nvcc $flags "$@" -c qecnetworks_b200/csrc/routing_fused2_wide.cu -o variants/obj/${name}_wide.o 2> variants/obj/${name}_wide.log &
wait
objs=$(ls qecnetworks_b200/build/*.o | grep -v "routing_fused2.o\|routing_fused2_wide.o")
nvcc -shared -o variants/libqec_$name.so variants/obj/$name.o variants/obj/${name}_wide.o $objs -gencode arch=compute_100a,code=sm_100a
grep -A2 "routing_fused2_bwd_kernelILi8ELb1" variants/obj/$name.log | grep -E "spill"
