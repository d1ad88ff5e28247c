"""The oracle against the DEFINITIONS of the published full-RNS CKKS construction, in Python big integers.

The reference's arithmetic lives in a closed backend with no golden vectors, so bit-exact parity is defined
against `oracle/` -- and the oracle itself has to be pinned to something that is not our own code.  This file pins it to
the mathematics, on a small ring (N = 2^8) where every definition can be evaluated directly, and again at N = 2^12 (the ring
of the emulation parity cases; the quadratic checks sampled there):

  * NTT          A[k] = a(psi^(2 bitrev(k) + 1)), inverse, and the negacyclic convolution theorem
  * Galois map   the NTT-domain gather realises m(X) -> m(X^g) mod X^N + 1 (S4)
  * conversion   fast base conversion = CRT lift + u D with 0 <= u < #sources; the exact variant = the centred lift (S5, S5')
  * rescale      (c - [c]_{q_l} centred) / q_l as an exact integer division of the CRT lift (S6)
  * key switch   c0' + c1' s  =  d s' + small   over Q_l, from the key equation evk = (-a s + e + P F s', a) (S6)
  * embedding    the encoded polynomial evaluated at the primitive roots zeta^(5^j) returns the slots (S9)
  * product      decrypt(ct1 * ct2) ~ z1 z2 and the rotation direction rotate(+r) = np.roll(+r) (App. A-6)

The engine is then tied to these definitions through the bit-exact engine-vs-oracle tests of test_engine_parity.py.
"""
from __future__ import annotations

import numpy as np
import pytest

from oracle.ckks_oracle import OracleCKKS, bitrev
from oracle.params import make_params

# the 256-coefficient ring (every check exhaustive) and N = 2^12, the ring of the emulation parity cases (the quadratic
# checks sampled: a sparse second factor, a subset of the slots)
@pytest.fixture(scope="module", params=[8, 12], ids=["n8", "n12"])
def orc(request):
    logn = request.param
    o = OracleCKKS(make_params(logn=logn, levels=4, dnum=2, hamming_weight=16 if logn == 8 else 64), seed=3)
    o.keygen_secret(); o.keygen_public(); o.keygen_relin()
    o.LOGN = logn
    return o


def rand_poly(o, idx, rng):
    return np.stack([rng.integers(0, o.moduli[i], 1 << o.LOGN, dtype=np.uint64) for i in idx])


def crt_lift(res, mods):
    """Residues [len(mods)] -> the unique integer in [0, prod mods)."""
    D = 1
    for m in mods:
        D *= m
    x = 0
    for r, m in zip(res, mods):
        h = D // m
        x += int(r) * h * pow(h % m, -1, m)
    return x % D, D


def centred(x, D):
    return x - D if x > D // 2 else x


def negacyclic_mul(a, b, q):
    N = len(a)
    out = [0] * N
    for j, bj in enumerate(b):
        if not bj:
            continue
        for i, ai in enumerate(a):
            k = i + j
            if k < N:
                out[k] = (out[k] + ai * bj) % q
            else:
                out[k - N] = (out[k - N] - ai * bj) % q
    return out


def test_ntt_is_evaluation_at_odd_powers_of_psi_and_a_ring_isomorphism(orc):
    LOGN = orc.LOGN; N = 1 << LOGN
    rng = np.random.default_rng(0)
    for i in (0, 1, orc.L + 1):                           # base prime, a scale prime, a special prime
        q = orc.moduli[i]
        psi = next(pow(x, (q - 1) // (2 * N), q) for x in range(2, 100) if pow(pow(x, (q - 1) // (2 * N), q), N, q) == q - 1)
        a, b = rand_poly(orc, [i], rng), rand_poly(orc, [i], rng)
        if N > 256:                                       # the schoolbook product is quadratic: a sparse second factor there
            b[0, rng.permutation(N)[: N - 9]] = 0
        A = orc.ntt(a, [i])
        for k in (0, 1, 2, 77, N - 1):                   # the definition, evaluated directly
            root = pow(psi, 2 * bitrev(k, LOGN) + 1, q)
            assert int(A[0, k]) == sum(int(c) * pow(root, j, q) for j, c in enumerate(a[0])) % q
        assert np.array_equal(orc.intt(A, [i]), a)
        prod = orc.intt(orc.mul(A, orc.ntt(b, [i]), [i]), [i])[0]
        assert [int(x) for x in prod] == negacyclic_mul([int(x) for x in a[0]], [int(x) for x in b[0]], q)


def test_galois_gather_is_the_substitution_x_to_x_g(orc):
    LOGN = orc.LOGN; N = 1 << LOGN
    rng = np.random.default_rng(1)
    i, q = 1, orc.moduli[1]
    a = rand_poly(orc, [i], rng)
    for g in (orc.galois_for_rotation(1), orc.galois_for_rotation(-7), orc.galois_conj()):
        want = [0] * N
        for j, c in enumerate(a[0]):                    # X^j -> X^(j g mod 2N), X^N = -1
            e = j * g % (2 * N)
            want[e % N] = (want[e % N] + (int(c) if e < N else -int(c))) % q
        got = orc.intt(orc.automorph(orc.ntt(a, [i]), g), [i])[0]
        assert [int(x) for x in got] == want


def test_base_conversion_against_the_crt_lift(orc):
    LOGN = orc.LOGN; N = 1 << LOGN
    rng = np.random.default_rng(2)
    src = [1, 2, 3]
    tgt = [0, 4] + list(range(orc.L + 1, orc.L + 1 + orc.K))
    x = rand_poly(orc, src, rng)
    fast = orc._baseconv(x, src, tgt)
    exact = orc._baseconv(x, src, tgt, exact=True)
    smods = [orc.moduli[i] for i in src]
    for k in range(0, N, 17):
        lift, D = crt_lift(x[:, k], smods)
        for r, t in enumerate(tgt):
            qt = orc.moduli[t]
            # fast conversion: the lift plus an overflow u D, 0 <= u < number of sources (S5)
            assert any((lift + u * D) % qt == int(fast[r, k]) for u in range(len(src)))
            # exact conversion: the centred representative (S5')
            assert centred(lift, D) % qt == int(exact[r, k])


def test_rescale_is_the_exact_division_of_the_crt_lift(orc):
    LOGN = orc.LOGN; N = 1 << LOGN
    rng = np.random.default_rng(3)
    level = 3
    idx = list(range(level + 1))
    c = rand_poly(orc, idx, rng)                                      # coefficient-domain residues
    out = orc.intt(orc.rescale_poly(orc.ntt(c, idx), level), idx[:-1])
    mods = [orc.moduli[i] for i in idx]
    ql = mods[-1]
    for k in range(0, N, 13):
        lift, Q = crt_lift(c[:, k], mods)
        d = int(c[level, k])
        d = d - ql if d > ql // 2 else d                               # centred [c]_{q_l}
        assert (lift - d) % ql == 0
        want = (lift - d) // ql
        for i in range(level):
            assert want % mods[i] == int(out[i, k])


def test_key_switch_satisfies_the_key_equation(orc):
    """Relinearisation key: c0' + c1' s = d s^2 + e over Q_l with |e| tiny against q (hybrid switching, S6)."""
    N = 1 << orc.LOGN
    rng = np.random.default_rng(4)
    level = orc.L
    idx = list(range(level + 1))
    d = rand_poly(orc, idx, rng)                                      # NTT-domain polynomial to switch
    c0, c1 = orc.key_switch(d, level, 0)
    s = orc.sk_ntt[idx]
    lhs = orc.add(c0, orc.mul(c1, s, idx), idx)
    rhs = orc.mul(d, orc.mul(s, s, idx), idx)
    err = orc.intt(orc.sub(lhs, rhs, idx), idx)
    mods = [orc.moduli[i] for i in idx]
    worst = 0
    for k in range(N):
        lift, Q = crt_lift(err[:, k], mods)
        worst = max(worst, abs(centred(lift, Q)))
    assert worst < 2 ** 30, worst                                      # Q is ~2^260: the switch is exact up to noise
    assert worst > 0                                                   # and it is a real (noisy) key switch


def test_embedding_is_evaluation_at_the_five_power_orbit(orc):
    LOGN = orc.LOGN; N = 1 << LOGN
    rng = np.random.default_rng(5)
    n = N // 2
    z = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    scale = 2.0 ** 40
    coef = orc.encode_coeffs(z, scale)                                 # N signed integers
    zeta = np.exp(1j * np.pi / N)                                      # primitive 2N-th root of unity
    e = 1
    check = set(range(n)) if n <= 128 else {0, 1, 2, 3, n // 2, n - 2, n - 1} | {int(x) for x in rng.integers(0, n, 9)}
    for j in range(n):                                                 # slot j <- m(zeta^(5^j)) / scale
        if j in check:
            val = sum(int(c) * zeta ** (e * t % (2 * N)) for t, c in enumerate(coef)) / scale
            assert abs(val - z[j]) < 1e-8
        e = e * 5 % (2 * N)


def test_product_rotation_and_conjugation_semantics(orc):
    LOGN = orc.LOGN; N = 1 << LOGN
    rng = np.random.default_rng(6)
    n = N // 2
    z1, z2 = (np.exp(2j * np.pi * rng.random(n)) for _ in range(2))
    c1, c2 = orc.encrypt(z1), orc.encrypt(z2)
    assert np.abs(orc.decrypt(orc.mul_ct(c1, c2)) - z1 * z2).max() < 1e-6
    assert np.abs(orc.decrypt(orc.rotate(c1, 3)) - np.roll(z1, 3)).max() < 1e-6      # positive = right (shift_rows.py:35-37)
    assert np.abs(orc.decrypt(orc.rotate(c1, -5)) - np.roll(z1, -5)).max() < 1e-6
    assert np.abs(orc.decrypt(orc.conjugate(c1)) - np.conj(z1)).max() < 1e-6


def test_bootstrap_oracle_refreshes_levels_within_tolerance():
    """oracle/bootstrap_oracle.py (ModRaise, CoeffToSlot, EvalMod, SlotToCoeff) on rings of 2^8 and 2^12
    coefficients: the refreshed ciphertext decrypts to the input within the stated tolerance (1e-3; the error grows
    with the ring: 5e-7 and 2e-5 here, 5.6e-4 measured on the engine at 2^16) and has at least 5 levels left."""
    from oracle.bootstrap_oracle import BootstrapOracle
    for logn, hw in ((8, 16), (12, 64)):
        o = OracleCKKS(make_params(logn=logn, levels=21, dnum=3, hamming_weight=hw, fresh_level=14), seed=2)
        o.keygen_secret(); o.keygen_public(); o.keygen_relin()
        B = BootstrapOracle(o, K=25, degree=47, double_angle=3)
        rng = np.random.default_rng(0)
        z = np.exp(2j * np.pi * rng.random(o.n))
        spent = o.level_down(o.encrypt(z), 1)                       # a ciphertext that has run out of levels
        out = B.bootstrap(spent)
        assert out.level >= 5 and out.level >= B.out_level
        assert np.abs(o.decrypt(out) - z).max() < 1e-3
