#!/bin/bash
# build a timing variant of the library with routing_fused3.cu compiled with extra -D flags (the other objects come from
# the regular build) -> variants/libqec_NAME.so:   scripts/build_variant3.sh NAME -DFOO ...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants/obj
flags="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -ftz=true -std=c++17 -Xcompiler -fPIC -Xptxas -v"
nvcc $flags "$@" -c qecnetworks_b200/csrc/routing_fused3.cu -o variants/obj/$name.o 2> variants/obj/$name.log
objs=$(ls qecnetworks_b200/build/*.o | grep -v "routing_fused3.o")
nvcc -shared -o variants/libqec_$name.so variants/obj/$name.o $objs -gencode arch=compute_100a,code=sm_100a
