"""Bit-exact parity of the engine against the integer oracle (oracle/ckks_oracle.py) under identical
parameters, keys and seed: parameters, keys, encrypt, decode, multiply, rotate, conjugate, constants,
level alignment, power basis, raw NTT / automorphism / key-switch.

The same test bodies run twice: on the emulation build of the CUDA sources (CPU, N = 2^12) and, marked
`gpu`, on the product library on a B200 (N = 2^12 and N = 2^16).
"""
from __future__ import annotations

import numpy as np
import pytest

import backend
from oracle.ckks_oracle import OracleCKKS
from oracle.params import make_params

CASES = [
    pytest.param(("emu", 12, 6), id="emu-n12"),
    pytest.param(("emu", 16, 5), id="emu-n16"),      # the two-round pass A (LOGR = 8) only exists at N = 2^16
    pytest.param(("cuda", 12, 6), id="cuda-n12", marks=pytest.mark.gpu),
    pytest.param(("cuda", 16, 5), id="cuda-n16", marks=pytest.mark.gpu),
    # the benchmark's own parameter set (bench.py): 22 ciphertext primes, 9 special primes below 2^50, dnum 3 (digits of 8
    # limbs, the last one ragged at the fresh level 14), Hamming weight 192 -- k_base_convert_fp<7/8/9/10>, the 9-prime
    # ModDown and the ragged ModUp are compared bit for bit here, not only through decoded bytes
    pytest.param(("cuda", 16, 21, 192, 14), id="cuda-n16-bench", marks=pytest.mark.gpu),
    # the round-1 chain: 61-bit special primes (integer-pipe NTT limbs, wide sources and integer targets in the basis
    # conversion: k_base_convert_fp<NS, 1> and the 128-bit multiply-accumulate branch)
    pytest.param(("emu", 12, 6, 64, -1, 5, 61, 60), id="emu-n12-p61"),
    pytest.param(("cuda", 16, 21, 192, 14, 5, 61, 60), id="cuda-n16-bench-p61", marks=pytest.mark.gpu),
    # the uniform chain of the first half of round 2: q_0 of 60 bits (integer-pipe limb 0, a wide source and an integer
    # target in every conversion, limb-0 decryption at any level).  The default is the descending-scale chain (q_0 next to
    # 2^50 as well, S_0 = 2^40, S_l -> 2^50 upwards; decryption aligns to level 0 first, spec S10)
    pytest.param(("emu", 12, 6, 64, -1, 5, None, 60), id="emu-n12-q60"),
    pytest.param(("cuda", 16, 21, 192, 14, 5, None, 60), id="cuda-n16-bench-q60", marks=pytest.mark.gpu),
]


class Pair:
    def __init__(self, which, logn, levels, hw=64, fresh=-1, seed=5, p_bits=None, q0_bits=None):
        mod = backend.use_emulation() if which == "emu" else backend.use_cuda()
        kw = dict(fresh_level=fresh) if fresh >= 0 else {}
        if p_bits is not None:
            kw["p_bits"] = p_bits
        if q0_bits is not None:
            kw["q0_bits"] = q0_bits
        self.eng = mod.Engine(logn=logn, levels=levels, dnum=3, hamming_weight=hw, seed=seed, **kw)
        self.params = make_params(logn=logn, levels=levels, dnum=3, hamming_weight=hw, **kw)
        self.orc = OracleCKKS(self.params, seed=seed)
        self.N = 1 << logn
        self.n = self.N // 2
        self.lib = self.eng._lib
        self.sk = self.eng.create_secret_key()
        self.pk = self.eng.create_public_key(self.sk)
        self.rk = self.eng.create_relinearization_key(self.sk)
        self.orc.keygen_secret()
        self.orc.keygen_public()
        self.orc.keygen_relin()

    def export(self, ct):
        a = np.zeros((ct.polynomial_count, ct.level + 1, self.N), dtype=np.uint64)
        assert self.lib.ckks_ct_export(self.eng._ptr, ct._h, a) == 0
        return a


@pytest.fixture(scope="module", params=CASES)
def pair(request):
    return Pair(*request.param)


def test_parameters_agree(pair):
    P = pair.eng.params()
    assert P["q"] == pair.params.q and P["p"] == pair.params.p
    assert P["scales"] == pair.params.scales
    assert P["alpha"] == pair.params.alpha
    assert all(x % (2 * pair.N) == 1 for x in P["q"] + P["p"])


def test_keys_bit_exact(pair):
    s = np.zeros(pair.N, dtype=np.int64)
    pair.lib.ckks_export_secret(pair.eng._ptr, s)
    assert np.array_equal(s, pair.orc.sk_coef)
    assert int(np.abs(s).sum()) == pair.eng.config["hamming_weight"]
    pk = np.zeros((2, pair.params.L + 1, pair.N), dtype=np.uint64)
    pair.lib.ckks_export_public(pair.eng._ptr, pk)
    assert np.array_equal(pk, pair.orc.pk)
    ev = np.zeros(pair.orc.evk[0].shape, dtype=np.uint64)
    pair.lib.ckks_export_switch_key(pair.eng._ptr, 0, ev)
    assert np.array_equal(ev, pair.orc.evk[0])


def test_raw_ntt_roundtrip_and_parity(pair):
    rng = np.random.default_rng(1)
    mods = list(range(len(pair.orc.moduli)))
    a = np.stack([rng.integers(0, m, pair.N, dtype=np.uint64) for m in pair.orc.moduli])
    b = a.copy()
    marr = np.asarray(mods, dtype=np.int32)
    assert pair.lib.ckks_test_ntt(pair.eng._ptr, b, len(mods), marr, 0) == 0
    assert np.array_equal(b, pair.orc.ntt(a, mods))
    assert pair.lib.ckks_test_ntt(pair.eng._ptr, b, len(mods), marr, 1) == 0
    assert np.array_equal(b, a)
    # edge values: 0, 1, q-1
    e = np.zeros_like(a)
    for i, m in enumerate(pair.orc.moduli):
        e[i, 0], e[i, 1], e[i, -1] = 1, m - 1, m - 1
    f = e.copy()
    pair.lib.ckks_test_ntt(pair.eng._ptr, f, len(mods), marr, 0)
    assert np.array_equal(f, pair.orc.ntt(e, mods))
    # largest magnitudes for the lazy (unreduced) intermediate values of both arithmetic paths (tools/ntt_fp_bounds.py)
    for pat in range(3):
        e = np.zeros_like(a)
        for i, m in enumerate(pair.orc.moduli):
            if pat == 0:
                e[i, :] = m - 1
            elif pat == 1:
                e[i, ::2] = m - 1
            else:
                e[i, :] = m // 2
        f = e.copy()
        pair.lib.ckks_test_ntt(pair.eng._ptr, f, len(mods), marr, 0)
        assert np.array_equal(f, pair.orc.ntt(e, mods))
        pair.lib.ckks_test_ntt(pair.eng._ptr, f, len(mods), marr, 1)
        assert np.array_equal(f, e)
        pair.lib.ckks_test_ntt(pair.eng._ptr, f, len(mods), marr, 1)      # inverse first, then forward
        pair.lib.ckks_test_ntt(pair.eng._ptr, f, len(mods), marr, 0)
        assert np.array_equal(f, e)


def test_raw_automorphism(pair):
    rng = np.random.default_rng(2)
    a = np.stack([rng.integers(0, m, pair.N, dtype=np.uint64) for m in pair.orc.moduli[:3]])
    for g in (pair.orc.galois_for_rotation(1), pair.orc.galois_for_rotation(-7), pair.orc.galois_conj()):
        b = a.copy()
        assert pair.lib.ckks_test_automorph(pair.eng._ptr, b, 3, g) == 0
        assert np.array_equal(b, pair.orc.automorph(a, g))
    assert pair.lib.ckks_galois_for_rotation(pair.eng._ptr, 5) == pair.orc.galois_for_rotation(5)
    assert pair.lib.ckks_galois_for_rotation(pair.eng._ptr, -5) == pair.orc.galois_for_rotation(-5)


@pytest.mark.parametrize("level_off", [0, 2])
def test_raw_key_switch(pair, level_off):
    rng = np.random.default_rng(3)
    level = pair.params.L - level_off
    poly = np.stack([rng.integers(0, pair.params.q[i], pair.N, dtype=np.uint64) for i in range(level + 1)])
    out = np.zeros((2, level + 1, pair.N), dtype=np.uint64)
    assert pair.lib.ckks_test_key_switch(pair.eng._ptr, poly, level, 0, out) == 0
    k0, k1 = pair.orc.key_switch(poly, level, 0)
    assert np.array_equal(out[0], k0) and np.array_equal(out[1], k1)


def test_encrypt_decrypt(pair):
    rng = np.random.default_rng(0)
    z = rng.normal(size=pair.n) + 1j * rng.normal(size=pair.n)
    ct, oc = pair.eng.encrypt(z), pair.orc.encrypt(z)
    assert np.array_equal(pair.export(ct), oc.c)
    d, od = pair.eng.decrypt(ct), pair.orc.decrypt(oc)
    assert np.array_equal(d, od)                      # fp64 embedding is bit-exact too (no FMA contraction)
    assert np.abs(d - z).max() < 1e-8


def test_homomorphic_ops_bit_exact(pair):
    rng = np.random.default_rng(4)
    eng, orc = pair.eng, pair.orc
    z1 = np.exp(2j * np.pi * rng.random(pair.n))
    z2 = np.exp(2j * np.pi * rng.random(pair.n))
    c1, c2 = eng.encrypt(z1), eng.encrypt(z2)
    o1, o2 = orc.encrypt(z1), orc.encrypt(z2)
    m, om = eng.multiply(c1, c2, pair.rk), orc.mul_ct(o1, o2)
    assert np.array_equal(pair.export(m), om.c)
    assert np.abs(eng.decrypt(m) - z1 * z2).max() < 1e-7
    for steps in (1, -3, pair.n // 4, 0):
        r, orr = eng.rotate(c1, None, steps), orc.rotate(o1, steps)
        assert np.array_equal(pair.export(r), orr.c)
        assert np.abs(eng.decrypt(r) - np.roll(z1, steps)).max() < 1e-7      # SURVEY App. A-6 direction
    cj, ocj = eng.conjugate(c1), orc.conjugate(o1)
    assert np.array_equal(pair.export(cj), ocj.c)
    assert np.abs(eng.decrypt(cj) - np.conj(z1)).max() < 1e-7
    k, ok = eng.multiply(c1, 0.25 - 95.47j), orc.mul_const(o1, 0.25 - 95.47j)
    assert np.array_equal(pair.export(k), ok.c)
    # level alignment: level L-1 product plus a fresh level-L ciphertext
    a, oa = eng.add(m, c1), orc.add_ct(om, o1)
    assert np.array_equal(pair.export(a), oa.c)
    s, os_ = eng.subtract(c2, m), orc.sub_ct(o2, om)
    assert np.array_equal(pair.export(s), os_.c)
    ac, oac = eng.add(c1, 1.0), orc.add_const(o1, 1.0)
    assert np.array_equal(pair.export(ac), oac.c)
    mask = (np.arange(pair.n) % 3 == 0).astype(np.complex128)
    mp, omp = eng.multiply(c1, eng.encode(mask)), orc.mul_plain_vec(o1, mask)
    assert np.array_equal(pair.export(mp), omp.c)
    assert np.abs(eng.decrypt(mp) - z1 * mask).max() < 1e-7
    z0 = eng.multiply(c1, 0.0)
    assert z0.level == c1.level - 1 and np.abs(eng.decrypt(z0)).max() < 1e-9


def test_power_basis_and_level_errors(pair):
    rng = np.random.default_rng(6)
    eng, orc = pair.eng, pair.orc
    z = np.exp(2j * np.pi * rng.random(pair.n))
    c, o = eng.encrypt(z), orc.encrypt(z)
    pb, opb = eng.make_power_basis(c, 8, pair.rk), orc.power_basis(o, 8)
    for x, y in zip(pb, opb):
        assert np.array_equal(pair.export(x), y.c)
    assert [x.level for x in pb] == [c.level - d for d in (0, 1, 2, 2, 3, 3, 3, 3)]
    assert np.abs(eng.decrypt(pb[7]) - z ** 8).max() < 1e-6
    # pruned basis (XOR4's odd exponents): same products, so bit-identical to the oracle's full basis; 6 and 8 are skipped
    c0 = eng.counters()
    sp = eng.make_power_basis_sparse(c, 8, [1, 3, 5, 7], pair.rk)
    assert eng.counters()["mul_cc"] - c0["mul_cc"] == 5
    assert [x is not None for x in sp] == [True, True, True, True, True, False, True, False]
    for k in (1, 2, 3, 4, 5, 7):
        assert np.array_equal(pair.export(sp[k - 1]), opb[k - 1].c)
    with pytest.raises(RuntimeError, match="out of range"):
        eng.make_power_basis_sparse(c, 8, [9], pair.rk)
    low = eng.level_down(c, 1)
    with pytest.raises(RuntimeError) as ei:
        eng.make_power_basis(low, 8, pair.rk)
    assert "level" in str(ei.value) and "positive" in str(ei.value)
    bottom = eng.level_down(c, 0)
    with pytest.raises(RuntimeError, match="positive"):
        eng.multiply(bottom, bottom, pair.rk)
    t3 = eng.multiply(c, c)                                   # no relin key: 3 polynomials
    assert t3.polynomial_count == 3
    r2 = eng.relinearize(t3, pair.rk)
    # on the descending-scale chain decryption aligns to level 0 (scale 2^40): the rescale rounding of the UN-relinearised
    # ciphertext carries s^2 (2e-7 at N = 2^16, Hamming weight 192); stated tolerance 1e-6 there, 1e-7 on the uniform chain
    tol3 = 1e-7 if eng.params()["q"][0] > 1 << 55 else 1e-6
    assert np.abs(eng.decrypt(r2) - z * z).max() < 1e-7 and np.abs(eng.decrypt(t3) - z * z).max() < tol3
    with pytest.raises(RuntimeError, match="should have 3 polynomials"):
        eng.relinearize(c, pair.rk)


def test_hoisted_rotations_decrypt(pair):
    rng = np.random.default_rng(8)
    z = np.exp(2j * np.pi * rng.random(pair.n))
    c = pair.eng.encrypt(z)
    pair.orc.encrypt(z)                      # keeps the two encryption counters (spec S8 stream ids) in step
    steps = [pair.n // 4, pair.n // 2, 3 * pair.n // 4, 0, -1]
    outs = pair.eng.rotate_many(c, None, steps)
    for s, r in zip(steps, outs):
        assert np.abs(pair.eng.decrypt(r) - np.roll(z, s)).max() < 1e-7


def test_hoisted_rotations_bit_exact(pair):
    rng = np.random.default_rng(9)
    z = np.exp(2j * np.pi * rng.random(pair.n))
    c, o = pair.eng.encrypt(z), pair.orc.encrypt(z)
    steps = [pair.n // 4, -2, 0]
    for got, want in zip(pair.eng.rotate_many(c, None, steps), pair.orc.rotate_hoisted(o, steps)):
        assert np.array_equal(pair.export(got), want.c)


def test_fused_lut2_bit_exact_and_equals_termwise(pair):
    """ckks_lut2 (one relinearisation) against the oracle restatement, and against the reference's term-by-term
    evaluation order (xor4_lut.py:63-74) on decrypted slots."""
    rng = np.random.default_rng(10)
    eng, orc = pair.eng, pair.orc
    za, zb = np.exp(2j * np.pi * rng.random(pair.n)), np.exp(2j * np.pi * rng.random(pair.n))
    ca, cb, oa, ob = eng.encrypt(za), eng.encrypt(zb), orc.encrypt(za), orc.encrypt(zb)
    pa, pb = eng.make_power_basis(ca, 4, pair.rk), eng.make_power_basis(cb, 4, pair.rk)
    opa, opb = orc.power_basis(oa, 4), orc.power_basis(ob, 4)
    A = [None] + pa + [eng.conjugate(pa[0])]
    B = [None] + pb + [eng.conjugate(pb[0])]
    OA = {i + 1: x for i, x in enumerate(opa)}
    OB = {i + 1: x for i, x in enumerate(opb)}
    OA[5], OB[5] = orc.conjugate(opa[0]), orc.conjugate(opb[0])
    terms = [(1, 1, 0.5 - 0.25j), (3, 2, -95.47j), (1, 4, 2.0), (5, 5, 1.0 + 1.0j), (2, 3, -0.125), (3, 1, 7.0)]
    got = eng.lut2(A, B, terms)
    want = orc.lut2(OA, OB, terms)
    assert got.level == want.level == min(x.level for x in pa + pb) - 2
    assert np.array_equal(pair.export(got), want.c)
    P = {1: za, 2: za ** 2, 3: za ** 3, 4: za ** 4, 5: np.conj(za)}
    Q = {1: zb, 2: zb ** 2, 3: zb ** 3, 4: zb ** 4, 5: np.conj(zb)}
    ref = sum(c * P[p] * Q[q] for p, q, c in terms)
    assert np.abs(eng.decrypt(got) - ref).max() < 1e-6
    # term-by-term, as the reference loop does
    acc = eng.subtract(A[1], A[1])
    for p, q, c in terms:
        acc = eng.add(acc, eng.multiply(eng.multiply(A[p], B[q], pair.rk), eng.encode(np.full(pair.n, c))))
    assert np.abs(eng.decrypt(acc) - eng.decrypt(got)).max() < 1e-6


def test_fused_lincomb_bit_exact(pair):
    rng = np.random.default_rng(11)
    eng, orc = pair.eng, pair.orc
    z = np.exp(2j * np.pi * rng.random(pair.n))
    c, o = eng.encrypt(z), orc.encrypt(z)
    pb, opb = eng.make_power_basis(c, 5, pair.rk), orc.power_basis(o, 5)      # levels L, L-1, L-2, L-2, L-3
    coef = np.array([0.5, -1.25 + 2j, 3.0j, 0.0625, -7.5 - 0.5j])
    got, want = eng.lincomb(pb, coef), orc.lincomb(opb, coef)
    assert np.array_equal(pair.export(got), want.c)
    assert got.level == pb[-1].level - 1
    assert np.abs(eng.decrypt(got) - sum(coef[k] * z ** (k + 1) for k in range(5))).max() < 1e-6


def test_device_renorm_equals_host_renorm(pair):
    """ckks_snap_zeta16 (decrypt, snap, re-encrypt on the device) == the reference's host renorm (pipeline.py:65-69)."""
    rng = np.random.default_rng(12)
    eng = pair.eng
    k = rng.integers(0, 16, pair.n)
    z = np.exp(-2j * np.pi * k / 16) * (1 + 0.05 * rng.normal(size=pair.n)) * np.exp(1j * 0.08 * rng.normal(size=pair.n))
    c = eng.encrypt(z)
    pair.orc.encrypt(z)
    out = eng.snap_zeta16(c, level=3)
    pair.orc._enc_counter += 1
    assert out.level == 3
    got = eng.decrypt(out)
    want = np.exp(-2j * np.pi * (np.rint(-np.angle(z) * 16 / (2 * np.pi)) % 16) / 16)
    assert np.abs(got - want).max() < 1e-8
    stride = pair.n // 16
    out2 = eng.snap_zeta16(c, level=-1, stride=stride)
    pair.orc._enc_counter += 1
    got2 = eng.decrypt(out2)
    want2 = np.ones(pair.n, dtype=np.complex128)
    want2[::stride] = want[::stride]
    assert np.abs(got2 - want2).max() < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("logn,levels", [(12, 6), (16, 5)])
def test_tensor_core_base_conversion_bit_exact(logn, levels):
    """The byte-sliced u8 tensor-core basis conversion (k_base_convert_mma, off by default: CKKS_BC_MMA=1) must give the
    same integers as the oracle's conversion in ModUp and in the exact ModDown: ct*ct, rotation, conjugation bit for bit."""
    import os
    os.environ["CKKS_BC_MMA"] = "1"
    try:
        p = Pair("cuda", logn, levels)
    finally:
        os.environ.pop("CKKS_BC_MMA", None)
    rng = np.random.default_rng(12)
    z1, z2 = (np.exp(2j * np.pi * rng.random(p.n)) for _ in range(2))
    c1, c2, o1, o2 = p.eng.encrypt(z1), p.eng.encrypt(z2), p.orc.encrypt(z1), p.orc.encrypt(z2)
    m, om = p.eng.multiply(c1, c2, p.rk), p.orc.mul_ct(o1, o2)
    assert np.array_equal(p.export(m), om.c)
    assert np.array_equal(p.export(p.eng.rotate(m, None, 7)), p.orc.rotate(om, 7).c)
    assert np.array_equal(p.export(p.eng.conjugate(c1)), p.orc.conjugate(o1).c)
    m2 = p.eng.multiply(m, m, p.rk)
    assert np.array_equal(p.export(m2), p.orc.mul_ct(om, om).c)


def test_edge_cases_of_the_reference_surface(pair):
    """The corners the reference's callers actually touch: short ("ragged") vectors are zero-padded,
    over-long ones rejected; rotations by 0 / slot_count / negative multiples are identities; a degree-1 power basis;
    operands many levels apart are aligned; constants as large as the XOR table's 95.5; the zero ciphertext; real input."""
    eng = pair.eng
    n = pair.n
    rng = np.random.default_rng(9)
    short = np.exp(2j * np.pi * rng.random(5))
    got = eng.decrypt(eng.encrypt(short))
    assert np.abs(got[:5] - short).max() < 1e-8 and np.abs(got[5:]).max() < 1e-8            # padded with zeros
    with pytest.raises(ValueError):
        eng.encrypt(np.ones(n + 1))
    with pytest.raises(ValueError):
        eng.encode(np.ones((2, n // 2)))
    z = np.exp(2j * np.pi * rng.random(n))
    c = eng.encrypt(z)
    for steps in (0, n, -n, 3 * n):
        assert np.abs(eng.decrypt(eng.rotate(c, None, steps)) - z).max() < 1e-8
    assert np.abs(eng.decrypt(eng.rotate(c, None, n + 2)) - np.roll(z, 2)).max() < 1e-7
    pb = eng.make_power_basis(c, 1, pair.rk)
    assert len(pb) == 1 and np.array_equal(pair.export(pb[0]), pair.export(c))
    deep = c
    for _ in range(c.level - 1):
        deep = eng.multiply(deep, 1.0)                                                       # one level per call
    assert deep.level == 1
    s = eng.add(c, deep)                                                                     # levels L and 1
    assert s.level == 1 and np.abs(eng.decrypt(s) - 2 * z).max() < 1e-6
    big = eng.multiply(c, 95.5)
    assert np.abs(eng.decrypt(big) - 95.5 * z).max() < 1e-5
    zero = eng.subtract(c, c)
    assert zero.level == c.level and np.abs(eng.decrypt(zero)).max() < 1e-9
    real = eng.decrypt(eng.encrypt(np.linspace(-1.0, 1.0, n)))                               # float64 input
    assert np.abs(real - np.linspace(-1.0, 1.0, n)).max() < 1e-8
    assert eng.decrypt(eng.encrypt(np.zeros(0))).shape == (n,)                                # empty input = all zeros


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["1", "2"])
def test_cluster_ntt_bit_exact(mode):
    """The single-kernel NTT (8-CTA cluster, DSMEM exchange; CKKS_NTT_CLUSTER=1: forward, =2: also the inverse and the fused
    forward variants; off by default -- measured slower than the two-pass kernels, profiles/README.md): keys, raw transforms,
    homomorphic operations and key switches with a RAGGED last digit (in-place transform of a batch whose slices have
    different item counts: ADVICE r1 -- surplus clusters must not store), repeated, stay bit-identical to the oracle."""
    import os
    os.environ["CKKS_NTT_CLUSTER"] = mode
    try:
        p = Pair("cuda", 16, 5)              # alpha = 2: levels 4, 2 have a ragged last digit
    finally:
        os.environ.pop("CKKS_NTT_CLUSTER", None)
    q = Pair("cuda", 12, 6)                  # the switch is per engine: an engine made afterwards is unaffected
    test_keys_bit_exact(p)
    test_raw_ntt_roundtrip_and_parity(p)
    test_homomorphic_ops_bit_exact(p)
    test_homomorphic_ops_bit_exact(q)
    rng = np.random.default_rng(3)
    for level in (4, 2):
        poly = np.stack([rng.integers(0, p.params.q[i], p.N, dtype=np.uint64) for i in range(level + 1)])
        k0, k1 = p.orc.key_switch(poly, level, 0)
        for _ in range(25):
            out = np.zeros((2, level + 1, p.N), dtype=np.uint64)
            assert p.lib.ckks_test_key_switch(p.eng._ptr, poly, level, 0, out) == 0
            assert np.array_equal(out[0], k0) and np.array_equal(out[1], k1)


@pytest.mark.parametrize("which,logn,levels", [("emu", 12, 6), pytest.param("cuda", 16, 5, marks=pytest.mark.gpu)])
def test_basis_conversion_paths_agree(which, logn, levels):
    """The FP64-pipe basis conversion (k_base_convert_fp, default) and the 128-bit integer one (CKKS_BC_FP=0) are two
    evaluations of the same exact sums: raw key switches at every level (every digit count, ragged last digits, q_0 as
    a split 60-bit source and as an integer-pipe target) give identical words, and both equal the oracle's."""
    import os
    fp = Pair(which, logn, levels)
    os.environ["CKKS_BC_FP"] = "0"
    try:
        integer = Pair(which, logn, levels)
    finally:
        os.environ.pop("CKKS_BC_FP", None)
    rng = np.random.default_rng(11)
    for level in range(levels, -1, -1):
        poly = np.stack([rng.integers(0, fp.params.q[i], fp.N, dtype=np.uint64) for i in range(level + 1)])
        want = fp.orc.key_switch(poly, level, 0)
        for p in (fp, integer):
            out = np.zeros((2, level + 1, p.N), dtype=np.uint64)
            assert p.lib.ckks_test_key_switch(p.eng._ptr, poly, level, 0, out) == 0
            assert np.array_equal(out[0], want[0]) and np.array_equal(out[1], want[1]), (which, level)


@pytest.mark.parametrize("which,logn,levels", [("emu", 12, 6), pytest.param("cuda", 16, 5, marks=pytest.mark.gpu)])
def test_level_alignment_on_the_rescale_epilogue(which, logn, levels):
    """Level alignment with the integer scalar riding on the rescale (k on the dropped limb after its inverse transform, k a
    inside the pass-B epilogue, NttFuse::ep_k; default) against k_mul_scalar on every limb followed by a rescale
    (CKKS_ALIGN_FUSE=0): identical words at every target level, for a 2- and a 3-polynomial ciphertext and for a batched
    handle, and equal to the oracle's alignment (spec S6)."""
    import os
    fused = Pair(which, logn, levels)
    os.environ["CKKS_ALIGN_FUSE"] = "0"
    try:
        plain = Pair(which, logn, levels)
    finally:
        os.environ.pop("CKKS_ALIGN_FUSE", None)
    rng = np.random.default_rng(17)
    z = np.exp(2j * np.pi * rng.random(fused.n))
    zb = np.exp(2j * np.pi * rng.random((3, fused.n)))
    oc = fused.orc.encrypt(z)
    made = []
    for p in (fused, plain):                                          # the first encryption of each engine = the oracle's first
        c = p.eng.encrypt(z)
        made.append((c, p.eng.multiply(c, c), p.eng.encrypt(zb)))     # a 3-polynomial product one level down, a handle of 3
    for target in range(levels - 1, -1, -1):
        words = []
        for p, (c, t3, b) in zip((fused, plain), made):
            lb = p.eng.level_down(b, target)
            nbw = np.zeros((3, 2, target + 1, p.N), dtype=np.uint64)
            assert p.lib.ckks_ct_export(p.eng._ptr, lb._h, nbw) == 0
            words.append((p.export(p.eng.level_down(c, target)),
                          p.export(p.eng.level_down(t3, target)) if target < levels - 1 else None, nbw))
        assert np.array_equal(words[0][0], words[1][0]), target
        assert words[0][1] is None or np.array_equal(words[0][1], words[1][1]), target
        assert np.array_equal(words[0][2], words[1][2]), target
        assert np.array_equal(words[0][0], fused.orc.level_down(oc, target).c), target
