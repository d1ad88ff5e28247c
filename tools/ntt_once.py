"""A few batched NTT calls (B x L limbs, N = 2^16) for ncu captures.  NTT_ONCE="L,B" (default 28,6).
L = 21 is one integer-path limb (q_0) + 20 FP64-path limbs; L = 28 adds the 7 special primes (integer path)."""
import ctypes as C, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "aes-implementation-fhe_b200")]
import desilofhe
L, B = (int(x) for x in os.environ.get("NTT_ONCE", "28,6").split(","))
eng = desilofhe.Engine(logn=16, levels=20, dnum=3, seed=1)
ms = C.c_float()
for inverse in (0, 1):
    desilofhe._capi.check(eng._lib.ckks_bench_ntt(eng._ptr, L, B, inverse, 3, C.byref(ms)))
    print("inverse" if inverse else "forward", L, B, ms.value, "ms", L * B * 2 * 65536 * 8 / ms.value / 1e6, "GB/s alg")
