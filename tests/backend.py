"""Backend selection for the test-suite.

`-m gpu` tests run the product library (nvcc, sm_100a) on a real device; everything else runs the
test-only emulation build of the same sources (tests/emu) so host orchestration and kernel arithmetic
are checked against the oracle without a GPU.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

# the engine-level tests check operations call for call against the oracle: deferred evaluation (desilofhe/lazy.py) is
# switched on explicitly by the tests that cover it
os.environ.setdefault("CKKS_B200_LAZY", "0")

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "aes-implementation-fhe_b200"), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def use_emulation():
    from emu.build import build
    os.environ["CKKS_B200_LIB"] = str(build())
    import desilofhe
    desilofhe._capi._lib = None
    return desilofhe


def use_cuda():
    os.environ.pop("CKKS_B200_LIB", None)
    import desilofhe
    desilofhe._capi._lib = None
    lib = desilofhe._capi.load()
    assert "cuda" in desilofhe._capi.backend(), "GPU tests must run the CUDA library"
    return desilofhe
