"""Wall-clock breakdown of one AES encryption round on the engine (sync after every phase), for guiding optimisation.
Prints JSON: seconds, key switches, limb-NTTs and kernel launches per phase."""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "aes-implementation-fhe_b200")]

import numpy as np


def main(reps: int = 3):
    import aes_fhe
    ctx = aes_fhe.EngineContext(1, mode="gpu", thread_count=1, logn=16, levels=21, fresh_level=14)
    eng = ctx.engine
    co = aes_fhe.load_all_coeffs()
    x4 = aes_fhe.XOR4LUT(ctx, co["xor4"])
    pipe = aes_fhe.AESPipeline(ctx, co, mixcolumns=aes_fhe.MixColFinal(ctx, x4),
                               inv_mixcolumns=aes_fhe.InvMixColumnsFHE(ctx, x4), use_hard_renorm_between_steps=True)
    drv = aes_fhe.FipsDriver(pipe, batched=True)
    stride = eng.slot_count // 16
    rng = np.random.default_rng(0)
    blocks = rng.integers(0, 256, (stride, 16), dtype=np.uint8)
    rks = aes_fhe.expand_aes128_key(np.arange(16, dtype=np.uint8))
    rk = pipe._prepare_round_keys([drv._perm(r) for r in rks])
    state = pipe.encoder.encode(drv._perm(blocks))
    acc = {}

    def phase(name, fn):
        eng.sync()
        c0, t0 = eng.counters(), time.perf_counter()
        out = fn()
        eng.sync()
        dt, c1 = time.perf_counter() - t0, eng.counters()
        a = acc.setdefault(name, {"s": 0.0, "keyswitch": 0, "ntt_limbs": 0, "launches": 0, "calls": 0})
        a["s"] += dt
        a["calls"] += 1
        for k in ("keyswitch", "ntt_limbs", "launches"):
            a[k] += c1[k] - c0[k]
        return out

    mix = pipe.mix
    for rep in range(reps + 1):
        if rep == 1:
            acc.clear()                     # first repetition is warm-up (lazy keys, tables)
        ct = phase("sub_bytes", lambda: pipe.sub_bytes(*state))
        ct = phase("renorm", lambda: pipe._renorm_pair(*ct))
        ct = phase("shift_rows", lambda: pipe.shift_rows(*ct))
        two, thr0 = phase("mix.gf2+gf3", lambda: mix._gf_shared([2, 3], *ct))       # one basis pair for both LUTs
        thr = phase("mix.rotations", lambda: mix._rot_pair(thr0, 1))
        steps = [-4 * k * mix.stride for k in (2, 3)]
        r2, r3 = phase("mix.rotations", lambda: list(zip(ctx.rotate_many(ct[0], steps), ctx.rotate_many(ct[1], steps))))
        a = phase("mix.xor", lambda: mix._xor_pair(two, thr))
        a = phase("renorm", lambda: mix._renorm_pair(*a))
        a = phase("mix.xor", lambda: mix._xor_pair(a, r2))
        a = phase("renorm", lambda: mix._renorm_pair(*a))
        a = phase("mix.xor", lambda: mix._xor_pair(a, r3))
        a = phase("renorm", lambda: mix._renorm_pair(*a))
        b0, b1 = phase("bootstrap", lambda: ctx.pair_map(ctx.bootstrap, (a[0],), (a[1],)))
        ct = phase("add_round_key", lambda: pipe.add_round_key(b0, b1, *rk[1]))
        ct = phase("renorm", lambda: pipe._renorm_pair(*ct))
    total = sum(v["s"] for v in acc.values())
    out = {k: {**{kk: (vv / reps) for kk, vv in v.items()}, "share": v["s"] / total} for k, v in acc.items()}
    out["total_s_per_round"] = total / reps
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
