"""Regenerate tests/golden/snap_reference_golden.json from the UNCHANGED reference noise-reduction classes
(`zeta16_noise_reducter.py`, `noise_reduction.py`, `snapper_1d_z16.py`) running on the slot stand-in.

Run in the build container only (needs /root/reference):  python tests/golden/make_snap_golden.py
Recorded per class: sha256 digest of the engine-call trace, op counts, and the output slots (256 complex values) for one
fixed noisy input -- what tests/test_snap.py pins the host mirror `aes_fhe/snap.py` to."""
import json
import os
import sys
from collections import Counter
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
os.environ.setdefault("STANDIN_SLOTS", "256")

import numpy as np  # noqa: E402

import refload  # noqa: E402
from oracle import slot_standin as ss  # noqa: E402


def noisy_codewords(n, seed=3, sigma=0.02):
    rng = np.random.default_rng(seed)
    z = np.exp(-2j * np.pi * rng.integers(0, 16, n) / 16)
    return z * (1 + sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n)))


def lut1d_coeffs():
    """A 1-D LUT on zeta16: the nibble map v -> (5 v + 3) mod 16 as a polynomial in x = zeta16^v (16-point DFT)."""
    w = np.exp(-2j * np.pi / 16)
    vals = np.array([w ** ((5 * v + 3) % 16) for v in range(16)])
    return np.fft.ifft(vals)              # f(w^v) = sum_k c_k w^(v k)  with  c_k = (1/16) sum_v f(w^v) w^(-v k)


def counts(eng):
    c = Counter()
    for (_, op), n in eng.counters.items():
        c[op] += n
    return dict(sorted(c.items()))


def main():
    ref = refload.load(ss)
    n = int(os.environ["STANDIN_SLOTS"])
    z = noisy_codewords(n)
    out = {"slots": n, "classes": {}}
    makers = {
        "Zeta16NoiseReducer": lambda ctx: ref.zeta16_noise_reducter.Zeta16NoiseReducer(ctx),
        "Zeta16SnapNoMul": lambda ctx: ref.zeta16_noise_reducter.Zeta16SnapNoMul(ctx),
        "Zeta16Snap": lambda ctx: ref.zeta16_noise_reducter.Zeta16Snap(ctx),
        "NoiseReducer": lambda ctx: ref.noise_reduction.NoiseReducer(ctx),
        "Zeta16Snap1D": lambda ctx: ref.snapper_1d_z16.Zeta16Snap1D(ctx, lut1d_coeffs()),
    }
    for name, make in makers.items():
        ctx = ref.engine_context.EngineContext(1, mode="cpu", thread_count=4)
        ctx.engine.trace_enabled = True
        obj = make(ctx)
        ct = ctx.encrypt(z)
        ctx.engine.reset_trace()
        y = ctx.decrypt(obj.apply(ct))
        out["classes"][name] = {"digest": ctx.engine.trace_digest(), "ops": counts(ctx.engine),
                                "re": [float(v) for v in y.real], "im": [float(v) for v in y.imag]}
    (HERE / "snap_reference_golden.json").write_text(json.dumps(out))
    print("wrote", HERE / "snap_reference_golden.json")


if __name__ == "__main__":
    main()
