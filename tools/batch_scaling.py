"""Key-switch / ct*ct / NTT throughput against the batch size of the handle (batch dimension) at N = 2^16,
bench parameters.  One JSON object; algorithmic bytes per key switch from SURVEY.md 8d."""
from __future__ import annotations

import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "aes-implementation-fhe_b200")]


def run(levels=21, dnum=3, iters=10, batches=(1, 2, 4, 8), lvls=(21, 14, 9, 5)) -> dict:
    import desilofhe
    eng = desilofhe.Engine(logn=16, levels=levels, dnum=dnum, seed=1)
    assert "cuda" in eng.backend
    sk = eng.create_secret_key(); eng.create_public_key(sk); eng.create_relinearization_key(sk)
    lib, ptr = eng._lib, eng._ptr
    P = eng.params()
    N, K = 1 << 16, len(P["p"])
    peak = 6556.2
    try:
        peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
    except Exception:
        pass
    ms = C.c_float()
    out = {"hbm_peak_gbs": peak, "alpha": P["alpha"], "K": K, "rotate": {}, "mul": {}, "ntt": {}}
    for lvl in lvls:
        beta = -(-(lvl + 1) // P["alpha"])
        ks_bytes = N * 8 * ((lvl + 1) + 2 * (lvl + 1))            # input poly + output ct, per item
        evk_bytes = N * 8 * 2 * beta * (lvl + 1 + K)              # the key: once per batch
        for nb in batches:
            desilofhe._capi.check(lib.ckks_bench_rotate_batch(ptr, lvl, nb, iters, C.byref(ms)))
            alg = nb * (ks_bytes + evk_bytes)                     # SURVEY 8d formula x items (key counted per item)
            out["rotate"][f"l{lvl}_b{nb}"] = {"ms_per_call": ms.value, "rot_per_s": nb * 1e3 / ms.value,
                                             "frac_hbm_survey_bytes": alg / (ms.value * 1e-3) / 1e9 / peak,
                                             "frac_hbm_key_once": (nb * ks_bytes + evk_bytes) / (ms.value * 1e-3) / 1e9 / peak}
            desilofhe._capi.check(lib.ckks_bench_mul_batch(ptr, lvl, nb, iters, C.byref(ms)))
            out["mul"][f"l{lvl}_b{nb}"] = {"ms_per_call": ms.value, "mul_per_s": nb * 1e3 / ms.value}
    for nl in (6, 15, 22, 29):
        for z in (1, 2, 4, 6):
            for inv in (0, 1):
                desilofhe._capi.check(lib.ckks_bench_ntt(ptr, nl, z, inv, iters, C.byref(ms)))
                limbs = nl * z
                out["ntt"][f"{'inv' if inv else 'fwd'}_{limbs}limbs"] = {"us": ms.value * 1e3, "us_per_limb": ms.value * 1e3 / limbs,
                                                                      "frac_hbm": limbs * 2 * N * 8 / (ms.value * 1e-3) / 1e9 / peak}
    return out


if __name__ == "__main__":
    print(json.dumps(run(), indent=1))
