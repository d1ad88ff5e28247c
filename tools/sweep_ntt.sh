#!/bin/bash
# builds the library with different NTT occupancy caps on the GPU box and runs the key-switch microbench for each
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 2 3 4 5; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DNTT_MIN_BLOCKS=$k \
      aes-implementation-fhe_b200/csrc/ckks_b200.cu -o aes-implementation-fhe_b200/lib/libckks_b200.so
  python tools/microbench.py > gpurun_out/microbench_ntt$k.json
  python - <<PY
import json
d=json.load(open("gpurun_out/microbench_ntt$k.json"))
print("NTT_MIN_BLOCKS=$k", {k:(round(v["ms"],4), round(v.get("alg_GBps",0))) for k,v in d.items() if isinstance(v,dict)})
PY
done
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared \
      aes-implementation-fhe_b200/csrc/ckks_b200.cu -o aes-implementation-fhe_b200/lib/libckks_b200.so
