#!/bin/bash
# Compile-time tuning variants of the product library for A/B runs on the GPU box:
#   tools/build_variants.sh name1:"-DX=1 -DY=2" name2:"-DZ=3" ...
# -> aes-implementation-fhe_b200/lib/variants/libckks_<name>.so (git-ignored, travels with gpurun);
# select one with CKKS_B200_LIB=<path> python bench.py ...
set -e
cd "$(dirname "$0")/.."
out=aes-implementation-fhe_b200/lib/variants
mkdir -p $out
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared $defs \
    aes-implementation-fhe_b200/csrc/ckks_b200.cu -o $out/libckks_$name.so &
done
wait
ls -la $out
