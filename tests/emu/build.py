"""Builds the CUDA-emulation library used by the `-m "not gpu"` tests (TEST INFRASTRUCTURE ONLY).

The very same sources as the product (`aes-implementation-fhe_b200/csrc/*.cu`) are compiled by g++
with -DCKKS_EMU: csrc/platform.cuh then maps the CUDA execution model onto plain loops (one CTA at a
time per OpenMP thread, threads of a CTA run region by region).  This lets the host orchestration and
the kernel arithmetic be checked against oracle/ in a container without a GPU.  The product never
loads this library (desilofhe/_capi.py looks only in aes-implementation-fhe_b200/lib unless a test
sets CKKS_B200_LIB), and bench.py / __graft_entry__.smoke() refuse to run on it.
"""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "aes-implementation-fhe_b200" / "csrc"
OUT = HERE / "libckks_emu.so"


def build(force: bool = False) -> Path:
    srcs = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "ckks_b200.h"]
    newest = max(p.stat().st_mtime for p in srcs)
    if OUT.exists() and OUT.stat().st_mtime >= newest and not force:
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off", "-DCKKS_EMU",
           "-x", "c++", str(CSRC / "ckks_b200.cu"), "-o", str(OUT)]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
