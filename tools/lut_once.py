"""One fused XOR4 (k_lut2), one S-box-style linear combination (k_lincomb) and one bootstrap (k_diag_mac_rows) on a batched
handle of 4 items at N = 2^16 on the default chain: the short command the round-2 ncu captures of the FP64-pipe
multiply-accumulate kernels are taken on (ncu --set full -k regex:"k_lut2|k_lincomb|k_diag_mac_rows")."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "aes-implementation-fhe_b200")]
import aes_fhe

ctx = aes_fhe.EngineContext(1, mode="gpu", thread_count=1, logn=16, levels=21, fresh_level=14, dnum=3, hamming_weight=192, seed=3)
eng = ctx.engine
rng = np.random.default_rng(0)
nib = rng.integers(0, 16, (4, eng.slot_count), dtype=np.uint8)
a, b = eng.encrypt_zeta16(nib, 5), eng.encrypt_zeta16(nib[::-1].copy(), 5)
x4 = aes_fhe.XOR4LUT(ctx, aes_fhe.load_all_coeffs()["xor4"])
out = x4.apply(a, b)
got = eng.decrypt_zeta16(eng.snap_zeta16(out, 5, 1))
assert np.array_equal(got, nib ^ nib[::-1]), "XOR4 bytes"
boot = ctx.bootstrap(eng.snap_zeta16(out, 0, 1))
assert np.array_equal(eng.decrypt_zeta16(boot), nib ^ nib[::-1]), "bootstrap bytes"
print("ok", boot.level)
