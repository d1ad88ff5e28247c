"""The UNCHANGED reference modules (`/root/reference/*.py`) running on the `desilofhe` drop-in.

Only possible where /root/reference exists (the build container); skipped on the GPU box.  The reference's own
`engine_context.py` is imported as is (tests/refload.py applies the two shims every backend needs),
its `EngineContext(signature=1, mode="cpu", thread_count=4)` constructor call of `test/test_aes_pipeline_roundtrip.py:116`
is issued verbatim, and the reference's XOR4LUT / AddRoundKey / SubBytes / ShiftRows run on the engine (emulation
build, ring shrunk through the CKKS_B200_ENGINE_OVERRIDES test hook)."""
from __future__ import annotations

import json
import os
import warnings

import numpy as np
import pytest

import backend
import refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="needs /root/reference")


@pytest.fixture(scope="module", params=["call-for-call", "deferred"])
def ref(request):
    """Both engine modes: every call executed at once, and the deferred evaluation of desilofhe/lazy.py (the default
    outside the tests) that turns the reference's term-by-term LUT loops into fused kernel calls."""
    mod = backend.use_emulation()
    os.environ["CKKS_B200_ENGINE_OVERRIDES"] = json.dumps({"logn": 12, "hamming_weight": 64,
                                                           "lazy": request.param == "deferred"})
    try:
        ns = refload.load(mod)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ns.ctx = ns.engine_context.EngineContext(signature=1, mode="cpu", thread_count=4)   # test/...roundtrip.py:116
        yield ns
    finally:
        os.environ.pop("CKKS_B200_ENGINE_OVERRIDES", None)


def test_reference_engine_context_surface(ref):
    ctx = ref.ctx
    assert ctx.engine.slot_count == 2048 and ctx.engine.slot_count % 16 == 0
    z = np.exp(-2j * np.pi * np.arange(2048) / 16)
    ct = ctx.encrypt(z)
    assert isinstance(ct, ref.engine_context.Ciphertext)
    assert np.abs(ctx.decrypt(ctx.multiply(ct, ct)) - z * z).max() < 1e-7               # ct*ct through engine_context.py:65-68
    assert np.abs(ctx.decrypt(ctx.rotate(ct, 128)) - np.roll(z, 128)).max() < 1e-7
    assert np.abs(ctx.decrypt(ctx.conjugate(ct)) - np.conj(z)).max() < 1e-7
    assert np.abs(ctx.decrypt(ctx.add_plain(ct, 1.0)) - (z + 1)).max() < 1e-7
    assert ctx.relinearize(ct) is ct                                                   # engine_context.py:139-145 ladder
    pb = ctx.make_power_basis_safe(ctx.engine.level_down(ct, 1), 8)                    # level ladder -> bootstrap -> retry
    assert len(pb) == 8 and ctx.bootstrap_stats()["count"] == 1
    assert np.abs(ctx.decrypt(pb[7]) - z ** 8).max() < 1e-2


def test_reference_ark_subbytes_shiftrows_bytes(ref):
    ctx = ref.ctx
    coeff_dir = ref.coeff_dir
    xor = ref.lut.load_coeff2d(coeff_dir / "xor4_coeffs.json", 16) if hasattr(ref.lut, "load_coeff2d") else None
    if xor is None:
        pytest.skip("reference loader name changed")
    x4 = ref.xor4_lut.XOR4LUT(ctx, xor)
    enc = ref.state_encoder.StateEncoder(ctx)
    key = np.frombuffer(bytes.fromhex("000102030405060708090a0b0c0d0e0f"), dtype=np.uint8)
    pt = np.frombuffer(bytes.fromhex("00112233445566778899aabbccddeeff"), dtype=np.uint8)
    ark = ref.add_round_key.AddRoundKey(x4)
    k0 = ctx.engine.counters()["keyswitch"]
    out = ark(*enc.encode(pt), *enc.encode(key))
    assert bytes(enc.decode(*out)).hex() == "00102030405060708090a0b0c0d0e0f0"          # golden enc.r0.ark
    ks = ctx.engine.counters()["keyswitch"] - k0
    assert ks == (2 * 19 if ctx.engine.lazy else 2 * 92), ks       # the unchanged XOR4 loop: 92 key switches, 19 when deferred
    sr = ref.shift_rows.ShiftRows(ctx)
    got = enc.decode(*sr.apply(*enc.encode(pt)))
    want = pt.reshape(4, 4).T.copy()                     # column-first packing: row r = bytes r + 4c (shift_rows.py:15-17)
    for r in range(4):
        want[r] = np.roll(want[r], -r)
    assert bytes(got) == bytes(want.T.reshape(-1))
    if ctx.engine.lazy:
        # the reference's SubBytes (sub_bytes_lut.py:46-73: 135 ct x ct, 134 conjugations call for call) through the deferred
        # engine: baby-step/giant-step polynomial evaluation, bytes against the S-box table embedded in the reference
        coeffs = refload.load_test_driver(ref).load_all_coeffs(coeff_dir)
        sb = ref.sub_bytes_lut.SubBytesLUT(ctx, coeffs["sub_hi"], coeffs["sub_lo"])
        k0 = ctx.engine.counters()["keyswitch"]
        got = enc.decode(*sb.apply(*enc.encode(pt)))
        ks = ctx.engine.counters()["keyswitch"] - k0
        sbox = np.array(ref.sub_bytes_lut.SBOX, dtype=np.uint8)            # the table embedded in the reference (sub_bytes_lut.py:86-103)
        assert bytes(got) == bytes(sbox[pt]), bytes(got).hex()
        assert ks < 60, ks

