#!/bin/bash
# usage: build_variant.sh NAME [nvcc -D flags...]   -> scratch/lib_NAME.so
name=$1; shift
mkdir -p scratch/obj_$name
for f in isla_capi isla_tree_tpt isla_tree_cta isla_rollout; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c islamcts_b200/csrc/$f.cu -o scratch/obj_$name/$f.o 2>/dev/null &
done; wait
/usr/local/cuda/bin/nvcc -shared -o scratch/lib_$name.so scratch/obj_$name/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
rm -rf scratch/obj_$name
