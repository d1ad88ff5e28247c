import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "aes-implementation-fhe_b200"
for p in (str(ROOT), str(PKG), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running")


# tests/_refscratch/ is the staged byte-for-byte copy of the reference's modules (tests/refload.py); it carries the
# reference's own test/ directory, which must not be collected as part of this suite.
collect_ignore_glob = ["_refscratch/*", "_refscratch/**/*"]
collect_ignore = ["_refscratch"]
