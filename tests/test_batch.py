"""The batch dimension of the engine (a handle holds nb independent ciphertexts, every operation runs all of
them through ONE set of kernel launches).  The contract: item i of a batched result is BIT-IDENTICAL to the operation
applied to item i alone, an nb = 1 operand is broadcast, and a batched encryption equals nb consecutive single ones.

Same bodies on the emulation build of the CUDA sources (CPU) and, marked `gpu`, on the product library on a B200
(N = 2^12 and the production ring N = 2^16 with a ragged last key-switch digit).
"""
from __future__ import annotations

import numpy as np
import pytest

import backend

CASES = [
    pytest.param(("emu", 12, 6, 3), id="emu-n12"),
    pytest.param(("cuda", 12, 6, 3), id="cuda-n12", marks=pytest.mark.gpu),
    pytest.param(("cuda", 16, 7, 3), id="cuda-n16", marks=pytest.mark.gpu),
]
NB = 3


class Env:
    def __init__(self, which, logn, levels, dnum):
        self.mod = backend.use_emulation() if which == "emu" else backend.use_cuda()
        self.kw = dict(logn=logn, levels=levels, dnum=dnum, hamming_weight=64, seed=17)
        self.eng = self.new_engine()
        rng = np.random.default_rng(4)
        n = self.eng.slot_count
        self.z = np.exp(2j * np.pi * rng.random((NB, n)))
        self.w = np.exp(2j * np.pi * rng.random((NB, n)))
        self.items_z = [self.eng.encrypt(v) for v in self.z]
        self.items_w = [self.eng.encrypt(v) for v in self.w]
        self.bz = self.eng.stack(self.items_z)
        self.bw = self.eng.stack(self.items_w)

    def new_engine(self):
        eng = self.mod.Engine(**self.kw)
        sk = eng.create_secret_key(); eng.create_public_key(sk)
        eng._rk = eng.create_relinearization_key(sk)
        return eng

    def raw(self, ct, eng=None):
        eng = eng or self.eng
        a = np.zeros((ct.batch, ct.polynomial_count, ct.level + 1, 2 * eng.slot_count), dtype=np.uint64)
        assert eng._lib.ckks_ct_export(eng._ptr, ct._h, a) == 0
        return a

    def same(self, batched, singles):
        """batched result == the per-item results, bit for bit"""
        got = self.raw(batched)
        assert batched.batch == len(singles)
        for i, s in enumerate(singles):
            want = self.raw(s)
            assert want.shape[0] == 1 and got[i].shape == want[0].shape
            assert np.array_equal(got[i], want[0]), f"item {i} differs"


@pytest.fixture(scope="module", params=CASES)
def env(request):
    return Env(*request.param)


def test_stack_unstack_and_decrypt(env):
    eng = env.eng
    assert env.bz.batch == NB and env.items_z[0].batch == 1
    env.same(env.bz, env.items_z)
    back = eng.unstack(env.bz)
    for a, b in zip(back, env.items_z):
        assert np.array_equal(env.raw(a), env.raw(b))
    d = eng.decrypt(env.bz)
    assert d.shape == (NB, eng.slot_count)
    for i in range(NB):
        assert np.array_equal(d[i], eng.decrypt(env.items_z[i]))
    with pytest.raises(RuntimeError):
        eng.add(env.bz, eng.stack(env.items_w[:2]))          # 3 items against 2: refused


def test_stack_of_batched_handles_and_slices(env):
    """ckks_ct_stack takes batched items (they contribute all their items, in order) and ckks_ct_slice cuts a range back out:
    how the two nibble planes of an AES state travel through ONE bootstrap (aes_fhe EngineContext.pair_apply)."""
    eng = env.eng
    both = eng.stack([env.bz, env.items_w[0], env.bw])
    assert both.batch == 2 * NB + 1
    env.same(both, env.items_z + [env.items_w[0]] + env.items_w)
    env.same(eng.batch_slice(both, 0, NB), env.items_z)
    env.same(eng.batch_slice(both, NB + 1, NB), env.items_w)
    env.same(eng.batch_slice(both, NB, 1), [env.items_w[0]])
    # an operation on the stack is the operation on every item (the contract that makes the stacked bootstrap exact)
    env.same(eng.batch_slice(eng.multiply(both, both, eng._rk), NB + 1, NB), [eng.multiply(c, c, eng._rk) for c in env.items_w])
    for bad in ((-1, 1), (0, 0), (2 * NB, 2)):
        with pytest.raises(RuntimeError):
            eng.batch_slice(both, *bad)
    with pytest.raises(RuntimeError):
        eng.stack([env.bz, eng.level_down(env.items_w[0], env.bz.level - 1)])      # shapes differ: refused


def test_batched_encryption_equals_consecutive_single_encryptions(env):
    a, b = env.new_engine(), env.new_engine()                  # same seed: same keys, encryption counters at 0
    cb = a.encrypt(env.z)
    singles = [b.encrypt(v) for v in env.z]
    got = env.raw(cb, a)
    for i, s in enumerate(singles):
        assert np.array_equal(got[i], env.raw(s, b)[0])
    nib = (np.arange(NB * a.slot_count, dtype=np.uint32) * 7 % 16).astype(np.uint8).reshape(NB, -1)
    cn = a.encrypt_zeta16(nib)
    sn = [b.encrypt_zeta16(r) for r in nib]
    got = env.raw(cn, a)
    for i, s in enumerate(sn):
        assert np.array_equal(got[i], env.raw(s, b)[0])
    assert np.array_equal(a.decrypt_zeta16(cn), nib)


def test_arithmetic_items_are_bit_identical(env):
    eng, rk = env.eng, env.eng._rk
    Z, W, bz, bw = env.items_z, env.items_w, env.bz, env.bw
    env.same(eng.add(bz, bw), [eng.add(a, b) for a, b in zip(Z, W)])
    env.same(eng.subtract(bz, bw), [eng.subtract(a, b) for a, b in zip(Z, W)])
    env.same(eng.negate(bz), [eng.negate(a) for a in Z])
    env.same(eng.multiply(bz, bw, rk), [eng.multiply(a, b, rk) for a, b in zip(Z, W)])
    env.same(eng.multiply(bz, 0.37 - 0.2j), [eng.multiply(a, 0.37 - 0.2j) for a in Z])
    env.same(eng.add(bz, 1.5), [eng.add(a, 1.5) for a in Z])
    mask = np.zeros(eng.slot_count); mask[::3] = 1.0
    pt = eng.encode(mask)
    env.same(eng.multiply(bz, pt), [eng.multiply(a, pt) for a in Z])
    env.same(eng.add(bz, pt), [eng.add(a, pt) for a in Z])
    env.same(eng.level_down(bz, 2), [eng.level_down(a, 2) for a in Z])
    # operands at different levels: the alignment runs batched as well
    lo = eng.multiply(bw, bw, rk)
    env.same(eng.add(bz, lo), [eng.add(a, eng.multiply(b, b, rk)) for a, b in zip(Z, W)])
    # three-polynomial path
    t3 = eng.multiply(bz, bw)
    s3 = [eng.multiply(a, b) for a, b in zip(Z, W)]
    env.same(t3, s3)
    env.same(eng.relinearize(t3), [eng.relinearize(x) for x in s3])
    env.same(eng.add(t3, bz), [eng.add(x, a) for x, a in zip(s3, Z)])
    env.same(eng.subtract(bz, t3), [eng.subtract(a, x) for x, a in zip(s3, Z)])


def test_broadcast_of_an_unbatched_operand(env):
    eng, rk = env.eng, env.eng._rk
    k = env.items_w[0]                                         # "round key": one ciphertext shared by the batch
    env.same(eng.add(env.bz, k), [eng.add(a, k) for a in env.items_z])
    env.same(eng.subtract(k, env.bz), [eng.subtract(k, a) for a in env.items_z])
    env.same(eng.multiply(env.bz, k, rk), [eng.multiply(a, k, rk) for a in env.items_z])
    env.same(eng.multiply(k, env.bz, rk), [eng.multiply(k, a, rk) for a in env.items_z])
    env.same(eng.multiply(k, env.bz), [eng.multiply(k, a) for a in env.items_z])


def test_galois_maps_and_power_basis(env):
    eng = env.eng
    env.same(eng.conjugate(env.bz), [eng.conjugate(a) for a in env.items_z])
    env.same(eng.rotate(env.bz, None, 5), [eng.rotate(a, None, 5) for a in env.items_z])
    low = eng.level_down(env.bz, 3)                            # ragged last key-switch digit
    env.same(eng.rotate(low, None, -9), [eng.rotate(eng.level_down(a, 3), None, -9) for a in env.items_z])
    steps = [1, 0, -3, 64]
    many = eng.rotate_many(env.bz, None, steps)
    for j, s in enumerate(steps):
        env.same(many[j], [eng.rotate_many(a, None, steps)[j] for a in env.items_z])
    pb = eng.make_power_basis(env.bz, 5)
    ps = [eng.make_power_basis(a, 5) for a in env.items_z]
    for k in range(5):
        env.same(pb[k], [p[k] for p in ps])
    d = eng.decrypt(pb[4])
    assert np.abs(d - env.z ** 5).max() < 1e-4


def test_fused_luts(env):
    eng = env.eng
    terms = [(1, 1, 0.5 + 0.25j), (1, 3, -0.125), (2, 1, 0.3j), (3, 3, 0.2), (3, 2, -0.4 + 0.1j)]

    def bases(x, y):
        A = [None] + eng.make_power_basis(x, 3) + [None] * 12
        B = [None] + eng.make_power_basis(y, 3) + [None] * 12
        return A, B

    A, B = bases(env.bz, env.bw)
    singles = [eng.lut2(*bases(a, b), terms) for a, b in zip(env.items_z, env.items_w)]
    env.same(eng.lut2(A, B, terms), singles)
    # the B basis unbatched (a round key's basis shared by all pairs)
    A1, B1 = bases(env.bz, env.items_w[1])
    env.same(eng.lut2(A1, B1, terms), [eng.lut2(*bases(a, env.items_w[1]), terms) for a in env.items_z])
    want = sum(c * env.z ** p * env.w ** q for p, q, c in terms)
    assert np.abs(eng.decrypt(eng.lut2(A, B, terms)) - want).max() < 1e-4
    # linear combination over several levels
    pz = eng.make_power_basis(env.bz, 4)
    co = [0.5, -0.25j, 0.125, 1.0 + 1.0j]
    env.same(eng.lincomb(pz, co), [eng.lincomb(eng.make_power_basis(a, 4), co) for a in env.items_z])


def test_device_renorm_of_a_batch(env):
    eng = env.eng
    rng = np.random.default_rng(9)
    nib = rng.integers(0, 16, (NB, eng.slot_count), dtype=np.uint8)
    ct = eng.encrypt_zeta16(nib)
    noisy = eng.multiply(ct, 1.001)                           # off the codewords, one level down
    snapped = eng.snap_zeta16(noisy, level=3, stride=1)
    assert snapped.batch == NB and snapped.level == 3
    assert np.array_equal(eng.decrypt_zeta16(snapped), nib)
    code = np.exp(-2j * np.pi * nib / 16)
    assert np.abs(eng.decrypt(snapped) - code).max() < 1e-6


def test_captured_graph_on_a_batch_draws_fresh_randomness_per_replay(env):
    """the encryptions inside a captured graph (device renorm) must not reuse the randomness frozen at capture
    time: two replays on the same input give different ciphertexts and equal decryptions."""
    eng, rk = env.eng, env.eng._rk
    rng = np.random.default_rng(10)
    nib = rng.integers(0, 16, (NB, eng.slot_count), dtype=np.uint8)
    ct = eng.encrypt_zeta16(nib)
    call = eng.capture(lambda a: [eng.snap_zeta16(eng.multiply(a, a, rk), level=2, stride=1)], [ct])
    first = env.raw(call(ct)[0]).copy()
    eng.sync()
    second = env.raw(call(ct)[0]).copy()
    assert not np.array_equal(first, second)
    want = (2 * nib.astype(np.int64)) % 16
    assert np.array_equal(eng.decrypt_zeta16(call.outputs[0]), want.astype(np.uint8))
    call.close()
