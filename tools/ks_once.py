"""A handful of rotations (key switches) at the top level of the N = 2^16 bootstrapping chain: the short command the ncu
captures in profiles/ are taken on."""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "aes-implementation-fhe_b200")]
import desilofhe

eng = desilofhe.Engine(logn=16, levels=21, dnum=3, seed=1)
sk = eng.create_secret_key(); eng.create_public_key(sk); eng.create_relinearization_key(sk)
ms = C.c_float()
for level in (21, 14):
    desilofhe._capi.check(eng._lib.ckks_bench_rotate(eng._ptr, level, 4, C.byref(ms)))
    print("rotate level", level, "ms", ms.value)
    desilofhe._capi.check(eng._lib.ckks_bench_mul(eng._ptr, level, 4, C.byref(ms)))
    print("mul level", level, "ms", ms.value)
