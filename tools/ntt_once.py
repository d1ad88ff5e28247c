"""A few batched NTT calls (6 x 28 limbs, N = 2^16) for ncu captures."""
import ctypes as C, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "aes-implementation-fhe_b200")]
import desilofhe
eng = desilofhe.Engine(logn=16, levels=20, dnum=3, seed=1)
ms = C.c_float()
for inverse in (0, 1):
    desilofhe._capi.check(eng._lib.ckks_bench_ntt(eng._ptr, 28, 6, inverse, 3, C.byref(ms)))
    print("inverse" if inverse else "forward", ms.value)
