"""A few batched rotations / multiplications (key switches on a handle of 8 items) at the fresh level of the N = 2^16
bootstrapping chain: the short command the round-2 ncu captures in profiles/ are taken on."""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "aes-implementation-fhe_b200")]
import desilofhe

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
eng = desilofhe.Engine(logn=16, levels=21, dnum=3, seed=1)
sk = eng.create_secret_key(); eng.create_public_key(sk); eng.create_relinearization_key(sk)
ms = C.c_float()
for level in (14, 21, 5):
    desilofhe._capi.check(eng._lib.ckks_bench_rotate_batch(eng._ptr, level, nb, 2, C.byref(ms)))
    print("rotate level", level, "batch", nb, "ms", ms.value)
    desilofhe._capi.check(eng._lib.ckks_bench_mul_batch(eng._ptr, level, nb, 2, C.byref(ms)))
    print("mul level", level, "batch", nb, "ms", ms.value)
