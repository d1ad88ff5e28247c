"""Bootstrapping precision at N = 2^16 on the device: max / mean slot error for random unit-modulus slots and for the
state-encoder layout (prints JSON)."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "aes-implementation-fhe_b200")]
import numpy as np
import desilofhe

eng = desilofhe.Engine(logn=16, levels=21, use_bootstrap=True, seed=1)
sk = eng.create_secret_key(); eng.create_public_key(sk); eng.create_relinearization_key(sk)
eng.create_conjugation_key(sk); eng.create_bootstrap_key(sk)
n = eng.slot_count
rng = np.random.default_rng(0)
out = {}
for name, z in (("random_unit", np.exp(2j * np.pi * rng.random(n))),
                ("zeta16_codewords", np.exp(-2j * np.pi * rng.integers(0, 16, n) / 16)),
                ("state_layout", np.where(np.arange(n) % (n // 16) == 0, np.exp(-2j * np.pi * (np.arange(n) // (n // 16)) / 16), 1.0))):
    errs = []
    for rep in range(3):
        got = eng.decrypt(eng.bootstrap(eng.encrypt(z)))
        errs.append(np.abs(got - z))
    e = np.concatenate(errs)
    out[name] = {"max": float(e.max()), "mean": float(e.mean()), "bits": float(-np.log2(e.max()))}
out["out_level"] = eng._lib.ckks_bootstrap_out_level(eng._ptr)
print(json.dumps(out, indent=1))
