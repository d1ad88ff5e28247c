"""Size-independent properties of the engine's operations (emulation build, N = 2^12): what must hold whatever the
parameters are -- additivity, commutativity (bit for bit: the operations are deterministic functions of the residues),
rotation composition, conjugation as an involution, distributivity within noise, level bookkeeping."""
from __future__ import annotations

import numpy as np
import pytest

import backend


@pytest.fixture(scope="module")
def env():
    mod = backend.use_emulation()
    eng = mod.Engine(logn=12, levels=6, dnum=3, hamming_weight=64, seed=13)
    sk = eng.create_secret_key(); eng.create_public_key(sk); rk = eng.create_relinearization_key(sk)
    rng = np.random.default_rng(2)
    zs = [np.exp(2j * np.pi * rng.random(eng.slot_count)) for _ in range(3)]
    return eng, rk, zs, [eng.encrypt(z) for z in zs]


def blob(eng, ct):
    return eng.serialize_ciphertext(ct)


def test_addition_and_multiplication_commute_bit_for_bit(env):
    eng, rk, zs, (a, b, c) = env
    assert blob(eng, eng.add(a, b)) == blob(eng, eng.add(b, a))
    assert blob(eng, eng.multiply(a, b, rk)) == blob(eng, eng.multiply(b, a, rk))
    assert blob(eng, eng.add(eng.add(a, b), c)) == blob(eng, eng.add(a, eng.add(b, c)))      # exact modular addition
    assert blob(eng, eng.subtract(eng.add(a, b), b)) == blob(eng, a)


def test_rotations_compose_and_conjugation_is_an_involution(env):
    eng, rk, zs, (a, b, c) = env
    n = eng.slot_count
    r = eng.rotate(eng.rotate(a, None, 5), None, 11)
    assert np.abs(eng.decrypt(r) - np.roll(zs[0], 16)).max() < 1e-6
    assert np.abs(eng.decrypt(r) - eng.decrypt(eng.rotate(a, None, 16))).max() < 1e-6
    back = eng.rotate(eng.rotate(a, None, 7), None, -7)
    assert np.abs(eng.decrypt(back) - zs[0]).max() < 1e-6
    assert np.abs(eng.decrypt(eng.conjugate(eng.conjugate(a))) - zs[0]).max() < 1e-6
    # rotation and conjugation commute; a rotation of a sum is the sum of the rotations (exactly: both are linear maps)
    assert np.abs(eng.decrypt(eng.conjugate(eng.rotate(a, None, 3))) - eng.decrypt(eng.rotate(eng.conjugate(a), None, 3))).max() < 1e-6
    assert np.abs(eng.decrypt(eng.rotate(eng.add(a, b), None, n // 4))
                  - eng.decrypt(eng.add(eng.rotate(a, None, n // 4), eng.rotate(b, None, n // 4)))).max() < 1e-6
    many = eng.rotate_many(a, None, [1, n // 2, -3])
    for ct, s in zip(many, (1, n // 2, -3)):
        assert np.abs(eng.decrypt(ct) - np.roll(zs[0], s)).max() < 1e-6


def test_multiplication_distributes_and_levels_are_consumed_one_at_a_time(env):
    eng, rk, zs, (a, b, c) = env
    lhs = eng.multiply(eng.add(a, b), c, rk)
    rhs = eng.add(eng.multiply(a, c, rk), eng.multiply(b, c, rk))
    assert lhs.level == rhs.level == a.level - 1
    assert np.abs(eng.decrypt(lhs) - eng.decrypt(rhs)).max() < 1e-6
    assert np.abs(eng.decrypt(lhs) - (zs[0] + zs[1]) * zs[2]).max() < 1e-6
    x = a
    for k in range(a.level):
        x = eng.multiply(x, a, rk)
        assert x.level == a.level - 1 - k
        assert np.abs(eng.decrypt(x) - zs[0] ** (k + 2)).max() < 1e-4
    with pytest.raises(RuntimeError, match="positive"):
        eng.multiply(x, a, rk)
    sq = eng.multiply(a, a, rk)
    pb = eng.make_power_basis(a, 2, rk)
    assert blob(eng, pb[1]) == blob(eng, sq)                          # make_power_basis(…, 2)[1] is exactly ct * ct
    lin = eng.lincomb([a, b, c], [0.5, -2.0 + 1j, 3.25j])
    assert np.abs(eng.decrypt(lin) - (0.5 * zs[0] + (-2.0 + 1j) * zs[1] + 3.25j * zs[2])).max() < 1e-6
