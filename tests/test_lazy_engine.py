"""Deferred evaluation behind the call-for-call API (desilofhe/lazy.py): the reference's term-by-term LUT loops
(xor4_lut.py:63-74, sub_bytes_lut.py:46-73, mixcol_final.py:80-99) issued against `Engine(lazy=True)` are evaluated by the
fused kernels -- same decoded values within rounding noise, same level bookkeeping, a fraction of the key switches --
and the reference's recovery ladders still see their RuntimeError at the call."""
from __future__ import annotations

import numpy as np
import pytest

import backend

CASES = [pytest.param("emu", id="emu-n12"), pytest.param("cuda", id="cuda-n12", marks=pytest.mark.gpu)]


@pytest.fixture(scope="module", params=CASES)
def engines(request):
    mod = backend.use_emulation() if request.param == "emu" else backend.use_cuda()
    out = []
    for lazy in (False, True):
        eng = mod.Engine(logn=12, levels=9, dnum=3, hamming_weight=64, seed=23, lazy=lazy)
        sk = eng.create_secret_key(); eng.create_public_key(sk)
        eng._rk = eng.create_relinearization_key(sk); eng._cj = eng.create_conjugation_key(sk)
        out.append(eng)
    return out


def reference_style_xor(eng, a, b, coeffs):
    """The call sequence of xor4_lut.py:27-74, verbatim in structure."""
    rk, cj = eng._rk, eng._cj

    def basis(ct):
        pos = eng.make_power_basis(ct, 8, rk)
        zero_like = eng.subtract(ct, ct)
        bs = {0: eng.add_plain(zero_like, 1.0)}
        for k in range(1, 9):
            bs[k] = pos[k - 1]
        for k in range(9, 16):
            bs[k] = eng.conjugate(pos[(16 - k) - 1], cj)
        return bs

    A, B = basis(a), basis(b)
    res = eng.subtract(A[0], A[0])
    for (p, q), c in coeffs.items():
        term = eng.multiply(A[p], B[q], rk)
        res = eng.add(res, eng.multiply(term, eng.encode(np.full(eng.slot_count, c, dtype=np.complex128))))
    return res


def test_xor4_loop_is_fused_and_agrees(engines):
    eager, lazy = engines
    import aes_fhe
    xor = aes_fhe.load_all_coeffs()["xor4"] / 256.0            # unit-modulus outputs
    coeffs = {(p, q): complex(xor[p, q]) for p in range(16) for q in range(16) if abs(xor[p, q]) > 1e-12}
    rng = np.random.default_rng(1)
    n = eager.slot_count
    na, nb = rng.integers(0, 16, n), rng.integers(0, 16, n)
    za, zb = np.exp(-2j * np.pi * na / 16), np.exp(-2j * np.pi * nb / 16)
    outs, ks = [], []
    for eng in (eager, lazy):
        a, b = eng.encrypt(za), eng.encrypt(zb)
        k0 = eng.counters()["keyswitch"]
        r = reference_style_xor(eng, a, b, coeffs)
        lvl = r.level                                          # known without evaluating
        val = eng.decrypt(r)
        ks.append(eng.counters()["keyswitch"] - k0)
        assert r.level == lvl == a.level - 5                   # 3 (basis) + 1 (ct x ct) + 1 (constant), SURVEY.md App. B
        outs.append(val)
    want = np.exp(-2j * np.pi * (na ^ nb) / 16)
    assert np.abs(outs[0] - want).max() < 1e-4 and np.abs(outs[1] - want).max() < 1e-4
    assert np.abs(outs[0] - outs[1]).max() < 1e-4
    assert ks[0] == 2 * (7 + 7) + 64                           # call for call: 78 ct x ct + 14 conjugations
    assert ks[1] == 2 * (5 + 4) + 1, ks                        # deferred: odd powers only, 4 conjugations per base, ONE relin


def test_linear_terms_and_conjugate_grouping(engines):
    """sub_bytes_lut.py:49-71 shape: sum_k c_k X_k over powers and conjugates of powers, plus a constant."""
    eager, lazy = engines
    rng = np.random.default_rng(2)
    n = eager.slot_count
    z = np.exp(2j * np.pi * rng.random(n))
    cs = {k: complex(rng.normal(), rng.normal()) / 16 for k in range(1, 16)}
    outs, ks = [], []
    for eng in (eager, lazy):
        ct = eng.encrypt(z)
        k0 = eng.counters()["keyswitch"]
        pos = eng.make_power_basis(ct, 8, eng._rk)
        res = eng.add_plain(eng.multiply(ct, 0.0), 0.25)
        for k in range(1, 16):
            bk = pos[k - 1] if k <= 8 else eng.conjugate(pos[16 - k - 1], eng._cj)
            res = eng.add(res, eng.multiply(bk, eng.encode(np.full(n, cs[k], dtype=np.complex128))))
        outs.append(eng.decrypt(res))
        ks.append(eng.counters()["keyswitch"] - k0)
        assert res.level == ct.level - 4
    want = 0.25 + sum(cs[k] * (z ** k if k <= 8 else np.conj(z ** (16 - k))) for k in range(1, 16))
    assert np.abs(outs[0] - want).max() < 1e-5 and np.abs(outs[1] - want).max() < 1e-5
    assert ks[0] == 7 + 7 and ks[1] == 7 + 1                   # one conjugation instead of seven


def test_errors_surface_at_the_call_and_values_are_shared(engines):
    _, lazy = engines
    z = np.exp(2j * np.pi * np.random.default_rng(3).random(lazy.slot_count))
    low = lazy.level_down(lazy.encrypt(z), 2)
    with pytest.raises(RuntimeError, match="level should be positive"):
        lazy.make_power_basis(low, 8, lazy._rk)                # xor4_lut.py:33-51 relies on the error coming NOW
    ct = lazy.level_down(lazy.encrypt(z), 4)
    pb = lazy.make_power_basis(ct, 4, lazy._rk)
    k0 = lazy.counters()["keyswitch"]
    again = lazy.make_power_basis(ct, 4, lazy._rk)            # mixcol_final.py:82-83 rebuilds identical bases
    v1, v2 = lazy.decrypt(pb[3]), lazy.decrypt(again[3])
    assert lazy.counters()["keyswitch"] - k0 == 2             # x^2 and x^4 once, not twice (x^3 never)
    assert np.array_equal(v1, v2) and np.abs(v1 - z ** 4).max() < 1e-5
    c1, c2 = lazy.conjugate(ct, lazy._cj), lazy.conjugate(ct, lazy._cj)
    assert c1 is c2
    # a deferred handle used by an operation outside the fused patterns is simply evaluated
    r = lazy.rotate(lazy.multiply(pb[1], pb[2], lazy._rk), None, 3)
    assert np.abs(lazy.decrypt(r) - np.roll(z ** 5, 3)).max() < 1e-4


def test_high_degree_polynomial_is_evaluated_baby_step_giant_step(engines):
    """sub_bytes_lut.py:60-71 shape: make_power_basis(b, 128) and sum_k c_k b^k, c_k conj(b^(256-k)) term by term.  Deferred:
    15 baby + 7 giant products, ONE relinearisation per partial sum and ONE conjugation, at the level the call-for-call
    evaluation reports."""
    eager, lazy = engines
    rng = np.random.default_rng(5)
    n = eager.slot_count
    byte = rng.integers(0, 256, n)
    z = np.exp(-2j * np.pi * byte / 256)
    table = rng.integers(0, 16, 256)                                          # some nibble-valued function of the byte
    coeffs = np.fft.ifft(np.exp(-2j * np.pi * table / 16))                     # f(z) = sum_k c_k z^k on the 256th roots
    outs, ks = [], []
    for eng in (eager, lazy):
        ct = eng.encrypt(z)
        k0 = eng.counters()["keyswitch"]
        pos = eng.make_power_basis(ct, 128, eng._rk)
        res = eng.add_plain(eng.multiply(ct, 0.0), complex(coeffs[0]))
        for k in range(1, 256):
            bk = pos[k - 1] if k <= 128 else eng.conjugate(pos[256 - k - 1], eng._cj)
            res = eng.add(res, eng.multiply(bk, eng.encode(np.full(n, coeffs[k], dtype=np.complex128))))
        assert res.level == ct.level - 8
        outs.append(eng.decrypt(res))
        assert res.level == ct.level - 8                                       # still true once it has a value
        ks.append(eng.counters()["keyswitch"] - k0)
    want = np.exp(-2j * np.pi * table[byte] / 16)
    assert np.abs(outs[0] - want).max() < 2e-3 and np.abs(outs[1] - want).max() < 2e-3
    assert ks[0] == 127 + 127 and ks[1] == 15 + 7 + 2 + 1, ks
