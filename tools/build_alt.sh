#!/bin/bash
# Kernel tuning experiments: build a variant of libgidcuda.so under build/alt/ (git-ignored, travels with gpurun).
# usage: tools/build_alt.sh name [extra nvcc flags...]   ->  build/alt/libgidcuda_<name>.so  (select with GIDCUDA_LIB)
set -e
name=$1; shift
mkdir -p build/alt
cd gid_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
  *.cu -o ../../build/alt/libgidcuda_$name.so
echo build/alt/libgidcuda_$name.so
