#!/bin/bash
# usage: build_variant.sh NAME [nvcc -D flags...]   -> profiles/experiments/lib_NAME.so
name=$1; shift
mkdir -p profiles/experiments/obj_$name
for f in isla_capi isla_tree_tpt isla_tree_cta isla_rollout; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c islamcts_b200/csrc/$f.cu -o profiles/experiments/obj_$name/$f.o 2>/dev/null &
done; wait
/usr/local/cuda/bin/nvcc -shared -o profiles/experiments/lib_$name.so profiles/experiments/obj_$name/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
rm -rf profiles/experiments/obj_$name
