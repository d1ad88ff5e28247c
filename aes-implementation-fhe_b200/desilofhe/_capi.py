"""ctypes binding of include/ckks_b200.h (libckks_b200.so, the sm_100a CUDA engine).

The product library lives in `aes-implementation-fhe_b200/lib/libckks_b200.so` and is built by
`__graft_entry__.build()` with nvcc.  If it is missing, or the process has no CUDA device, loading
fails loudly -- there is no CPU fallback.  `CKKS_B200_LIB` may point at another build of the same
ABI; the only other build that exists is the test-only CUDA-emulation library under `tests/emu/`,
which the `-m "not gpu"` tests select explicitly (bench.py and smoke() refuse it).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent.parent
DEFAULT_LIB = _PKG / "lib" / "libckks_b200.so"

OK, ERR_LEVEL, ERR_FORM, ERR_POLYS, ERR_OTHER = range(5)

_lib = None
_lib_path = None


class EngineLibraryMissing(ImportError):
    pass


def lib_path() -> Path:
    return Path(os.environ.get("CKKS_B200_LIB", str(DEFAULT_LIB)))


def load():
    global _lib, _lib_path
    path = lib_path()
    if _lib is not None and _lib_path == path:
        return _lib
    if not path.exists():
        raise EngineLibraryMissing(
            f"{path} not found: build the CUDA engine first (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU fallback.")
    L = C.CDLL(str(path))
    vp, i32, u64, dbl, lng = C.c_void_p, C.c_int, C.c_uint64, C.c_double, C.c_long
    pp = C.POINTER(vp)
    dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
    i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
    i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    lngp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")   # C long == int64 on this ABI
    sig = {
        "ckks_last_error": (C.c_char_p, []),
        "ckks_backend": (C.c_char_p, []),
        "ckks_launch_count": (lng, []),
        "ckks_engine_create_default": (i32, [i32] * 10 + [u64, i32, pp]),
        "ckks_engine_create": (i32, [i32, u64p, i32, u64p, i32, i32, i32, i32, i32, u64, i32, pp]),
        "ckks_engine_destroy": (None, [vp]),
        "ckks_sync": (i32, [vp]),
        "ckks_fork": (i32, [vp, i32]), "ckks_set_lane": (i32, [vp, i32]), "ckks_set_lanes_enabled": (i32, [vp, i32]), "ckks_join": (i32, [vp]),
        "ckks_slot_count": (i32, [vp]),
        "ckks_get_params": (i32, [vp] + [C.POINTER(i32)] * 5 + [vp, vp, vp]),
        "ckks_keygen_secret": (i32, [vp]), "ckks_keygen_public": (i32, [vp]), "ckks_keygen_relin": (i32, [vp]),
        "ckks_keygen_conjugation": (i32, [vp]), "ckks_keygen_rotation": (i32, [vp, lngp, i32]),
        "ckks_keygen_bootstrap": (i32, [vp]),
        "ckks_set_keys_external": (i32, [vp, i32]),
        "ckks_switch_key_ids": (i32, [vp, u64p, i32, C.POINTER(i32)]),
        "ckks_switch_key_buffer": (i32, [vp, u64, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "ckks_set_bootstrap_params": (i32, [vp, i32, i32, i32, i32, i32]),
        "ckks_encode": (i32, [vp, dp, i32, pp]),
        "ckks_encrypt": (i32, [vp, dp, i32, pp]),
        "ckks_decrypt": (i32, [vp, vp, dp]),
        "ckks_snap_zeta16": (i32, [vp, vp, i32, i32, pp]),
        "ckks_encrypt_zeta16": (i32, [vp, np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS"), i32, pp]),
        "ckks_decrypt_zeta16": (i32, [vp, vp, np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")]),
        "ckks_ct_batch": (i32, [vp]),
        "ckks_encrypt_batch": (i32, [vp, dp, i32, i32, pp]),
        "ckks_encrypt_zeta16_batch": (i32, [vp, np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS"), i32, i32, pp]),
        "ckks_ct_stack": (i32, [vp, pp, i32, pp]),
        "ckks_ct_item": (i32, [vp, vp, i32, pp]),
        "ckks_ct_slice": (i32, [vp, vp, i32, i32, pp]),
        "ckks_ct_free": (None, [vp, vp]), "ckks_pt_free": (None, [vp, vp]),
        "ckks_ct_level": (i32, [vp]), "ckks_ct_npoly": (i32, [vp]), "ckks_pt_level": (i32, [vp]),
        "ckks_add": (i32, [vp, vp, vp, pp]), "ckks_sub": (i32, [vp, vp, vp, pp]),
        "ckks_negate": (i32, [vp, vp, pp]),
        "ckks_mul": (i32, [vp, vp, vp, pp]), "ckks_mul_norelin": (i32, [vp, vp, vp, pp]),
        "ckks_relinearize": (i32, [vp, vp, pp]),
        "ckks_mul_const": (i32, [vp, vp, dbl, dbl, pp]), "ckks_mul_plain": (i32, [vp, vp, vp, pp]),
        "ckks_add_const": (i32, [vp, vp, dbl, dbl, pp]), "ckks_add_plain": (i32, [vp, vp, vp, pp]),
        "ckks_mul_i": (i32, [vp, vp, i32, pp]),
        "ckks_level_down": (i32, [vp, vp, i32, pp]),
        "ckks_power_basis": (i32, [vp, vp, i32, pp]),
        "ckks_power_basis_sparse": (i32, [vp, vp, i32, i32p, i32, pp]),
        "ckks_conjugate": (i32, [vp, vp, pp]),
        "ckks_rotate": (i32, [vp, vp, lng, pp]),
        "ckks_rotate_hoisted": (i32, [vp, vp, lngp, i32, pp]),
        "ckks_lut2": (i32, [vp, pp, pp, i32, i32p, i32p, dp, i32, pp]),
        "ckks_lincomb": (i32, [vp, pp, i32, dp, pp]),
        "ckks_bootstrap": (i32, [vp, vp, pp]),
        "ckks_bootstrap_out_level": (i32, [vp]),
        "ckks_counters": (i32, [vp, lngp]),
        "ckks_arena_stats": (i32, [vp, C.POINTER(lng), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
        "ckks_ct_export": (i32, [vp, vp, u64p]),
        "ckks_ct_import": (i32, [vp, i32, i32, u64p, pp]),
        "ckks_ct_import_batch": (i32, [vp, i32, i32, i32, u64p, pp]),
        "ckks_pt_export": (i32, [vp, vp, u64p]),
        "ckks_export_secret": (i32, [vp, i64p]),
        "ckks_export_public": (i32, [vp, u64p]),
        "ckks_export_switch_key": (i32, [vp, u64, u64p]),
        "ckks_test_ntt": (i32, [vp, u64p, i32, i32p, i32]),
        "ckks_test_automorph": (i32, [vp, u64p, i32, u64]),
        "ckks_test_key_switch": (i32, [vp, u64p, i32, u64, u64p]),
        "ckks_galois_for_rotation": (u64, [vp, lng]),
        "ckks_graph_create": (i32, [vp, C.POINTER(i32)]),
        "ckks_graph_enter": (i32, [vp, i32]), "ckks_graph_leave": (i32, [vp]),
        "ckks_graph_capture_begin": (i32, [vp, i32]), "ckks_graph_capture_end": (i32, [vp, i32]),
        "ckks_graph_capture_abort": (i32, [vp]),
        "ckks_graph_launch": (i32, [vp, i32, i32]), "ckks_graph_wait": (i32, [vp, i32]),
        "ckks_graph_destroy": (i32, [vp, i32]),
        "ckks_graph_info": (i32, [vp, i32, C.POINTER(lng), C.POINTER(lng), C.POINTER(C.c_size_t), C.POINTER(lng)]),
        "ckks_ct_assign": (i32, [vp, vp, vp, i32]),
        "ckks_ct_clear_memo": (i32, [vp, vp]),
        "ckks_timer_start": (i32, [vp]),
        "ckks_timer_stop_ms": (i32, [vp, C.POINTER(C.c_float)]),
        "ckks_profile_ntt_begin": (i32, [vp]),
        "ckks_profile_ntt_end": (i32, [vp, C.POINTER(dbl), C.POINTER(lng), C.POINTER(lng)]),
        "ckks_bench_ntt": (i32, [vp, i32, i32, i32, i32, C.POINTER(C.c_float)]),
        "ckks_bench_rotate": (i32, [vp, i32, i32, C.POINTER(C.c_float)]),
        "ckks_bench_rotate_lanes": (i32, [vp, i32, i32, i32, C.POINTER(C.c_float)]),
        "ckks_bench_mul": (i32, [vp, i32, i32, C.POINTER(C.c_float)]),
        "ckks_bench_rotate_batch": (i32, [vp, i32, i32, i32, C.POINTER(C.c_float)]),
        "ckks_bench_mul_batch": (i32, [vp, i32, i32, i32, C.POINTER(C.c_float)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)        # AttributeError here = the library does not export the ABI
        f.restype, f.argtypes = res, args
    L._symbols = list(sig)
    _lib, _lib_path = L, path
    return L


def backend() -> str:
    return load().ckks_backend().decode()


def is_emulation() -> bool:
    return "emulation" in backend()


def check(status: int):
    """Map a C status to the exception the reference's callers expect (engine_context.py:139-145,184-195)."""
    if status == OK:
        return
    msg = load().ckks_last_error().decode(errors="replace")
    if status == ERR_LEVEL:
        raise RuntimeError(f"ciphertext level should be positive: {msg}")
    if status == ERR_FORM:
        raise RuntimeError(f"operand is not in NTT form: {msg}")
    if status == ERR_POLYS:
        raise RuntimeError(f"ciphertext should have 3 polynomials: {msg}")
    raise RuntimeError(msg)
