#!/bin/bash
# rebuilds the library with host-side launch timing and reports how the host time of one bootstrap splits
set -e
cd "$(dirname "$0")/.."
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DCKKS_TIME_LAUNCHES \
    aes-implementation-fhe_b200/csrc/ckks_b200.cu -o aes-implementation-fhe_b200/lib/libckks_b200.so
python - <<'PY'
import sys, time, ctypes as C
sys.path[:0] = [".", "aes-implementation-fhe_b200"]
import numpy as np, desilofhe
eng = desilofhe.Engine(logn=16, levels=21, use_bootstrap=True, seed=1)
sk = eng.create_secret_key(); eng.create_public_key(sk); rk = eng.create_relinearization_key(sk)
eng.create_conjugation_key(sk); eng.create_bootstrap_key(sk)
lib = eng._lib; lib.ckks_launch_host_ms.restype = C.c_double
z = np.exp(2j*np.pi*np.random.default_rng(0).random(eng.slot_count))
a = eng.encrypt(z)
for name, fn, reps in (("bootstrap", lambda: eng.bootstrap(a), 3), ("mul", lambda: eng.multiply(a, a, rk), 50), ("rotate", lambda: eng.rotate(a, None, 7), 50)):
    fn(); eng.sync()
    l0, m0, t0 = lib.ckks_launch_count(), lib.ckks_launch_host_ms(), time.perf_counter()
    for _ in range(reps): out = fn()
    t1 = time.perf_counter(); eng.sync()
    n = lib.ckks_launch_count() - l0
    print(f"{name}: host {1e3*(t1-t0)/reps:.3f} ms/call, launches {n/reps:.0f}, inside cudaLaunch {(lib.ckks_launch_host_ms()-m0)/reps:.3f} ms ({1e3*(lib.ckks_launch_host_ms()-m0)/n:.2f} us each)")
PY
