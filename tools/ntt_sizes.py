"""Forward / inverse NTT timing over call sizes (limbs per call) at N = 2^16, bench chain: one JSON object."""
import ctypes as C, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "aes-implementation-fhe_b200")]
import desilofhe
eng = desilofhe.Engine(logn=16, levels=21, dnum=3, seed=1)
ms = C.c_float()
out = {}
for nl, z in ((15, 4), (24, 6), (22, 8), (31, 6)):
    for inv in (0, 1):
        desilofhe._capi.check(eng._lib.ckks_bench_ntt(eng._ptr, nl, z, inv, 20, C.byref(ms)))
        out[f"{'inv' if inv else 'fwd'}_{nl * z}"] = round(ms.value * 1e3, 2)
print(json.dumps(out))
