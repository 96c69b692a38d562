#!/bin/bash
# Debug build of libgidcuda.so with device-side assertions on every coefficient address and table
# index of the Huffman kernels (-DGIDK_CHECKED), IN PLACE of the normal library, for one run of the
# GPU test suite (compute-sanitizer is not available on the test pool):
#   tools/checked_build.sh && gpurun -- 'python -m pytest tests -m gpu -x -q' ; python -m gid_b200.build --force
set -e
cd "$(dirname "$0")/../gid_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DGIDK_CHECKED \
  *.cu -o ../lib/libgidcuda.so
echo "checked build in gid_b200/lib/libgidcuda.so -- rebuild with: python -m gid_b200.build --force"
