"""The C ABI boundary (include/ckks_b200.h): the product library and the test-only emulation build export every symbol
the header declares; the Python loader fails loudly when the CUDA library is missing (no CPU fallback); the
`--impl reference` arm of bench.py prints the contract's JSON line."""
from __future__ import annotations

import ctypes
import json
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "ckks_b200.h"
PRODUCT = ROOT / "aes-implementation-fhe_b200" / "lib" / "libckks_b200.so"


def declared():
    return sorted(set(re.findall(r"\b(ckks_[a-z0-9_]+)\s*\(", HEADER.read_text())))


def test_header_declares_the_reference_surface():
    names = declared()
    for must in ("ckks_encrypt", "ckks_decrypt", "ckks_encode", "ckks_mul", "ckks_add", "ckks_sub", "ckks_power_basis",
                 "ckks_conjugate", "ckks_rotate", "ckks_relinearize", "ckks_bootstrap", "ckks_keygen_secret",
                 "ckks_keygen_public", "ckks_keygen_relin", "ckks_keygen_conjugation", "ckks_keygen_rotation",
                 "ckks_keygen_bootstrap", "ckks_slot_count", "ckks_last_error"):
        assert must in names
    # every entry point that replaces a reference method cites the reference file it replaces
    txt = HEADER.read_text()
    assert txt.count("engine_context.py") >= 10


def test_product_library_exports_every_declared_symbol():
    if not PRODUCT.exists():
        import __graft_entry__ as g
        g.build_cuda()
    lib = ctypes.CDLL(str(PRODUCT))                     # loading needs no GPU; no compute call is made here
    missing = [n for n in declared() if not hasattr(lib, n)]
    assert not missing, missing
    lib.ckks_backend.restype = ctypes.c_char_p
    assert lib.ckks_backend() == b"cuda-sm_100a"


def test_emulation_library_exports_the_same_abi():
    from emu.build import build
    lib = ctypes.CDLL(str(build()))
    assert not [n for n in declared() if not hasattr(lib, n)]
    lib.ckks_backend.restype = ctypes.c_char_p
    assert b"emulation" in lib.ckks_backend()


def test_binding_covers_the_header_and_fails_loudly_without_the_library(tmp_path, monkeypatch):
    import backend
    mod = backend.use_emulation()
    bound = set(mod._capi.load()._symbols)
    # the ctypes binding declares a prototype for everything in the header except pure test/bench helpers it never calls
    unbound = set(declared()) - bound
    assert unbound <= {"ckks_engine_create", "ckks_pt_level", "ckks_launch_host_ms"}, unbound
    monkeypatch.setenv("CKKS_B200_LIB", str(tmp_path / "nope.so"))
    mod._capi._lib = None
    with pytest.raises(ImportError, match="no CPU fallback"):
        mod._capi.load()
    monkeypatch.delenv("CKKS_B200_LIB")
    mod._capi._lib = None


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="8")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "aes128_fhe_blocks_per_s" and line["higher_is_better"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"] > 0


@pytest.mark.parametrize("which", [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)])
def test_ciphertext_wire_format_roundtrip(which):
    """serialize -> deserialize gives a ciphertext that decrypts identically and multiplies bit-identically; a blob from
    another modulus chain, a truncated blob and an out-of-range residue are rejected (on the emulation
    build and on the B200)."""
    import numpy as np
    import backend
    mod = backend.use_emulation() if which == "emu" else backend.use_cuda()
    eng = mod.Engine(logn=12, levels=4, dnum=2, hamming_weight=32, seed=9)
    sk = eng.create_secret_key(); eng.create_public_key(sk); rk = eng.create_relinearization_key(sk)
    rng = np.random.default_rng(1)
    z = np.exp(2j * np.pi * rng.random(eng.slot_count))
    ct = eng.multiply(eng.encrypt(z), eng.encrypt(z), rk)
    blob = eng.serialize_ciphertext(ct)
    assert len(blob) == 28 + 2 * (ct.level + 1) * 4096 * 8
    back = eng.deserialize_ciphertext(blob)
    assert back.level == ct.level and back.polynomial_count == 2
    assert np.array_equal(eng.decrypt(back), eng.decrypt(ct))
    assert eng.serialize_ciphertext(eng.multiply(back, back, rk)) == eng.serialize_ciphertext(eng.multiply(ct, ct, rk))
    other = mod.Engine(logn=12, levels=3, dnum=2, hamming_weight=32, seed=9)
    with pytest.raises(ValueError, match="different ring or modulus chain"):
        other.deserialize_ciphertext(blob)
    with pytest.raises(ValueError, match="corrupt"):
        eng.deserialize_ciphertext(blob[:-8])
    bad = bytearray(blob); bad[28:36] = (2 ** 64 - 1).to_bytes(8, "little")
    with pytest.raises(ValueError, match="out of range"):
        eng.deserialize_ciphertext(bytes(bad))


def test_capture_rejects_host_round_trips_and_recovers():
    """A function with a hidden host round trip (decrypt) cannot be captured: the call raises before the driver sees it,
    the capture is abandoned, and the engine keeps working eagerly and can capture again; replays are bit-identical to
    the eager result on new inputs."""
    import numpy as np
    import backend
    mod = backend.use_emulation()
    eng = mod.Engine(logn=12, levels=4, dnum=2, hamming_weight=32, seed=4)
    sk = eng.create_secret_key(); eng.create_public_key(sk); rk = eng.create_relinearization_key(sk)
    rng = np.random.default_rng(0)
    z1, z2 = (np.exp(2j * np.pi * rng.random(eng.slot_count)) for _ in range(2))
    a, b = eng.encrypt(z1), eng.encrypt(z2)

    def bad(x, y):
        p = eng.multiply(x, y, rk)
        eng.decrypt(p)                       # device-to-host copy + synchronisation
        return [p]

    with pytest.raises(RuntimeError, match="not allowed while a graph is being captured"):
        eng.capture(bad, [a, b])
    assert np.abs(eng.decrypt(eng.multiply(a, b, rk)) - z1 * z2).max() < 1e-7        # eager path intact

    def good(x, y):
        return [eng.add(eng.multiply(x, y, rk), eng.rotate(x, None, 5))]

    call = eng.capture(good, [a, b])
    assert call.info()["capture_misses"] == 0 and call.info()["launches"] > 10
    out = call(b, a)[0]                                                                 # swapped inputs
    want = good(b, a)[0]
    assert eng.serialize_ciphertext(out) == eng.serialize_ciphertext(want)
    assert np.abs(eng.decrypt(out) - (z1 * z2 + np.roll(z2, 5))).max() < 1e-7
    with pytest.raises(ValueError):
        call(a)                                                                         # wrong number of inputs
    with pytest.raises(RuntimeError, match="shape mismatch"):
        call(eng.level_down(a, 2), b)                                                   # wrong level
    call.close()
    assert np.abs(eng.decrypt(eng.multiply(a, b, rk)) - z1 * z2).max() < 1e-7


def test_bench_line_contract_on_the_emulation_build():
    """bench.py's own arm, run end to end on the test-only emulation build (N = 2^12, marked invalid as a measurement):
    the JSON line must carry every key the driver and the judge read, the captured-graph path must be the one that ran,
    and the decrypted bytes must equal the plain FIPS-197 round."""
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--dry-run-emulation", "--pairs", "1", "--steps", "1",
                          "--warmup", "0", "--no-dec"], capture_output=True, text=True, timeout=1500, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "aes128_fhe_blocks_per_s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["dtype"] == "u64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and d["config"]["pairs_per_gpu"] == 1
    graphs = d["config"]["cuda_graphs"]                     # first round, middle round (replayed 9 times), last round
    assert set(graphs) == {"enc_first", "enc", "enc_last"}
    assert graphs["enc"]["launches"] > 1000 and all(g["capture_misses"] == 0 for g in graphs.values())
    assert d["gpu_launches"] >= graphs["enc_first"]["launches"] + 9 * graphs["enc"]["launches"] + graphs["enc_last"]["launches"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert d["bytes_exact_vs_fips197"] is True
    assert "invalid" in d                                   # a dry run never passes for a measurement
