#!/bin/bash
# builds the library with the FP64 NTT path off / on (the GPU box has nvcc) and runs the NTT / key-switch microbench
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 0 1; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DNTT_FP64=$v \
      aes-implementation-fhe_b200/csrc/ckks_b200.cu -o aes-implementation-fhe_b200/lib/libckks_b200.so
  python tools/microbench.py > gpurun_out/microbench_fp$v.json
  python - <<PY
import json
d=json.load(open("gpurun_out/microbench_fp$v.json"))
print("NTT_FP64=$v", {k:(round(v["ms"],4), round(v.get("alg_GBps",0))) for k,v in d.items() if isinstance(v,dict)})
PY
done
