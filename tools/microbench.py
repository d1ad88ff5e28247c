"""Kernel-level micro-benchmarks on resident random data: limb-NTT throughput against the HBM
roofline, rotations/s (key-switch) and ct*ct multiplications/s at N = 2^16.  Prints one JSON object."""
from __future__ import annotations

import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "aes-implementation-fhe_b200")]


def run(levels: int = 20, dnum: int = 3, iters: int = 20) -> dict:
    import desilofhe
    eng = desilofhe.Engine(logn=16, levels=levels, dnum=dnum, seed=1)
    assert "cuda" in eng.backend
    sk = eng.create_secret_key(); eng.create_public_key(sk); eng.create_relinearization_key(sk)
    lib, ptr = eng._lib, eng._ptr
    P = eng.params()
    N, nq, K = 1 << 16, len(P["q"]), len(P["p"])
    peak = 6556.2
    try:
        peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
    except Exception:
        pass
    ms = C.c_float()
    out = {"logn": 16, "levels": levels, "dnum": dnum, "K": K, "alpha": P["alpha"], "hbm_peak_gbs": peak}
    for inverse in (0, 1):
        for batches in (1, 6):
            desilofhe._capi.check(lib.ckks_bench_ntt(ptr, nq + K, batches, inverse, iters, C.byref(ms)))
            limbs = (nq + K) * batches
            gbs = limbs * 2 * N * 8 / (ms.value * 1e-3) / 1e9
            out[f"{'intt' if inverse else 'ntt'}_b{batches}"] = {"limbs": limbs, "ms": ms.value, "alg_GBps": gbs,
                                                                "frac_of_measured_hbm": gbs / peak}
    for level in (levels, levels // 2, 5):
        desilofhe._capi.check(lib.ckks_bench_rotate(ptr, level, iters, C.byref(ms)))
        beta = -(-(level + 1) // P["alpha"])
        alg = N * 8 * ((level + 1) + 2 * beta * (level + 1 + K) + 2 * (level + 1))
        out[f"rotate_l{level}"] = {"ms": ms.value, "per_s": 1e3 / ms.value, "alg_MiB": alg / 2 ** 20,
                                   "alg_GBps": alg / (ms.value * 1e-3) / 1e9,
                                   "frac_of_measured_hbm": alg / (ms.value * 1e-3) / 1e9 / peak}
        desilofhe._capi.check(lib.ckks_bench_mul(ptr, level, iters, C.byref(ms)))
        out[f"mul_l{level}"] = {"ms": ms.value, "per_s": 1e3 / ms.value}
    # element-wise: ct + ct at the top level (reads 2, writes 1 ciphertext of 2 x (levels+1) limbs), enqueued back to back
    import numpy as np
    rng = np.random.default_rng(0)
    a = eng.encrypt(np.exp(2j * np.pi * rng.random(eng.slot_count)), level=levels)
    b = eng.encrypt(np.exp(2j * np.pi * rng.random(eng.slot_count)), level=levels)
    for _ in range(5):
        eng.add(a, b)
    eng.sync()
    desilofhe._capi.check(lib.ckks_timer_start(ptr))
    reps = 200
    for _ in range(reps):
        eng.add(a, b)
    desilofhe._capi.check(lib.ckks_timer_stop_ms(ptr, C.byref(ms)))
    byts = 3 * 2 * nq * N * 8
    out["add_ct_ct_top"] = {"ms": ms.value / reps, "alg_MiB": byts / 2 ** 20, "alg_GBps": byts / (ms.value / reps * 1e-3) / 1e9,
                            "frac_of_measured_hbm": byts / (ms.value / reps * 1e-3) / 1e9 / peak,
                            "note": "includes the Python/ctypes enqueue of every call"}
    out["launches"] = eng.counters()["launches"]
    return out


if __name__ == "__main__":
    print(json.dumps(run(), indent=1))
