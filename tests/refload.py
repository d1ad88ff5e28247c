"""Load the UNCHANGED reference modules from /root/reference on an injected backend.

Only usable in the build container (the GPU box has no /root/reference).  Implements
the R0 shims of SURVEY.md Appendix C without editing any reference file: a scratch
directory of symlinks to the reference .py files, `generator/coeffs` -> `gen/coeff`
(H2), the `SubBytesLUT` alias (H1) and a `desilofhe` module object of our choosing.
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types
from pathlib import Path

REF = Path(os.environ.get("AESFHE_REFERENCE", "/root/reference"))
REF_MODULES = ["engine_context", "utils", "state_encoder", "lut", "xor4_lut", "add_round_key", "sub_bytes_lut",
               "shift_rows", "inv_shiftrows", "zeta16_noise_reducter", "noise_reduction", "snapper_1d_z16",
               "mixcol_final", "invmixcolumns_fhe", "pipeline"]


def available() -> bool:
    return (REF / "pipeline.py").exists()


def load(backend_module) -> types.SimpleNamespace:
    """Import the reference modules against `backend_module` (anything exposing Engine/Ciphertext)."""
    scratch = Path(tempfile.mkdtemp(prefix="refmods_"))
    for f in REF.glob("*.py"):
        (scratch / f.name).symlink_to(f)
    (scratch / "generator").mkdir()
    (scratch / "generator" / "coeffs").symlink_to(REF / "gen" / "coeff")
    for m in REF_MODULES + ["desilofhe"]:
        sys.modules.pop(m, None)
    shim = types.ModuleType("desilofhe")
    shim.Engine = backend_module.Engine
    shim.Ciphertext = backend_module.Ciphertext
    sys.modules["desilofhe"] = shim
    sys.path.insert(0, str(scratch))
    try:
        sb = importlib.import_module("sub_bytes_lut")
        sb.SubBytesLUT = sb.SubBytesLUTFastCached            # H1
        mods = {m: importlib.import_module(m) for m in REF_MODULES}
    finally:
        sys.path.remove(str(scratch))
        for m in REF_MODULES + ["desilofhe"]:
            sys.modules.pop(m, None)
    ns = types.SimpleNamespace(**mods)
    ns.coeff_dir = scratch / "generator" / "coeffs"
    return ns
