"""The AES-on-CKKS host stack on the real engine arithmetic (BASELINE.json configs 1, 2, 3, 5).

Byte-level truth: FIPS-197 (via `cryptography` AES-ECB and the plain round model of bench.py) and the
reference-semantics slot stand-in (oracle/slot_standin.py, pinned to the unchanged reference modules by
tests/golden).  CPU cases run the emulation build at N = 2^12; `gpu` cases run the product library at N = 2^16.
"""
from __future__ import annotations

import numpy as np
import pytest
from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

import backend
import aes_fhe
from bench import plain_round
from oracle import slot_standin as ss

CASES = [
    pytest.param(("emu", 12, 64), id="emu-n12"),
    pytest.param(("cuda", 16, 192), id="cuda-n16", marks=pytest.mark.gpu),
]


def ecb(key: bytes, data: bytes) -> bytes:
    e = Cipher(algorithms.AES(key), modes.ECB()).encryptor()
    return e.update(data) + e.finalize()


def make_pipe(ctx):
    co = aes_fhe.load_all_coeffs()
    x4 = aes_fhe.XOR4LUT(ctx, co["xor4"])
    return aes_fhe.AESPipeline(ctx, co, mixcolumns=aes_fhe.MixColFinal(ctx, x4),
                               inv_mixcolumns=aes_fhe.InvMixColumnsFHE(ctx, x4), use_hard_renorm_between_steps=True)


@pytest.fixture(scope="module", params=CASES)
def boot_ctx(request):
    which, logn, hw = request.param
    mod = backend.use_emulation() if which == "emu" else backend.use_cuda()
    ctx = aes_fhe.EngineContext(1, mode="gpu", thread_count=1, backend=mod, logn=logn, levels=21, fresh_level=14,
                                hamming_weight=hw, seed=1)
    return which, ctx


def test_bootstrap_precision_and_levels(boot_ctx):
    which, ctx = boot_ctx
    eng = ctx.engine
    n = eng.slot_count
    rng = np.random.default_rng(0)
    z = np.exp(2j * np.pi * rng.random(n))
    out = ctx.bootstrap(ctx.to_intt(ctx.encrypt(z)))
    assert out.level == eng._lib.ckks_bootstrap_out_level(eng._ptr) >= 5         # >= 5 after bootstrap
    err = np.abs(ctx.decrypt(out) - z).max()
    assert err < 1e-3, err                  # stated tolerance: 10 bits; measured 5.7e-4 at N = 2^16
    # the state encoding: ones everywhere, codewords on the stride grid
    v = np.ones(n, dtype=np.complex128)
    v[:: n // 16] = np.exp(-2j * np.pi * np.arange(16) / 16)
    err2 = np.abs(ctx.decrypt(ctx.bootstrap(ctx.encrypt(v))) - v).max()
    assert err2 < 1e-3, err2
    assert ctx.bootstrap_stats()["count"] == 2


def test_bootstrap_fused_inner_sums_are_bit_identical(boot_ctx, monkeypatch):
    """All inner sums of a BSGS matrix in one launch (k_diag_mac_rows, every baby rotation read once) against one launch
    per giant row (CKKS_DIAG_ROWS=0): the same ciphertext bit for bit, single and batched."""
    which, ctx = boot_ctx
    eng = ctx.engine
    n = eng.slot_count
    rng = np.random.default_rng(5)

    def export(ct):
        a = np.zeros((eng._lib.ckks_ct_batch(ct._h), ct.polynomial_count, ct.level + 1, 2 * n), dtype=np.uint64)
        assert eng._lib.ckks_ct_export(eng._ptr, ct._h, a) == 0
        return a

    cts = [ctx.to_intt(ctx.encrypt(np.exp(2j * np.pi * rng.random(n))))]
    cts.append(eng.encrypt(np.exp(2j * np.pi * rng.random((3, n))), level=5))
    for ct in cts:
        monkeypatch.setenv("CKKS_DIAG_ROWS", "0")
        per_row = export(ctx.bootstrap(ct))
        monkeypatch.delenv("CKKS_DIAG_ROWS")
        assert np.array_equal(export(ctx.bootstrap(ct)), per_row)


@pytest.mark.slow
def test_engine_bootstrap_agrees_with_the_bootstrap_oracle(boot_ctx):
    """The same bootstrapping spec evaluated by the engine (fused BSGS sums, hoisted and double-hoisted
    rotations, lanes) and by oracle/bootstrap_oracle.py (one operation at a time) under the same parameters and keys:
    decrypted slots agree within the stated tolerance 1e-3 (measured 2e-5 at N = 2^12), same output level plan."""
    which, ctx = boot_ctx
    from oracle.bootstrap_oracle import BootstrapOracle
    from oracle.ckks_oracle import OracleCKKS
    from oracle.params import make_params
    eng = ctx.engine
    logn, hw = (12, 64) if which == "emu" else (16, 192)        # on the B200: the production ring (the oracle takes minutes)
    orc = OracleCKKS(make_params(logn=logn, levels=21, dnum=3, hamming_weight=hw, fresh_level=14), seed=1)
    orc.keygen_secret(); orc.keygen_public(); orc.keygen_relin()
    assert eng.params()["q"] == [int(x) for x in orc.q] and eng.params()["p"] == [int(x) for x in orc.p]
    B = BootstrapOracle(orc, K=25, degree=47, double_angle=3)
    assert B.out_level == eng._lib.ckks_bootstrap_out_level(eng._ptr)
    rng = np.random.default_rng(31)
    z = np.exp(2j * np.pi * rng.random(eng.slot_count))
    got = ctx.decrypt(ctx.bootstrap(ctx.to_intt(ctx.encrypt(z))))
    want = orc.decrypt(B.bootstrap(orc.encrypt(z)))
    assert np.abs(got - want).max() < 1e-3 and np.abs(got - z).max() < 1e-3 and np.abs(want - z).max() < 1e-3


@pytest.mark.parametrize("fused", [False, True], ids=["call-for-call", "fused"])
def test_config1_ark_subbytes_fips_vector(boot_ctx, fused):
    """configs[0]: AddRoundKey + SubBytes on the FIPS-197 C.1 state, once issuing the reference's
    engine calls one for one and once through the fused LUT entry points."""
    which, ctx = boot_ctx
    ctx.fused = fused
    pipe = make_pipe(ctx)
    key = np.frombuffer(bytes.fromhex("000102030405060708090a0b0c0d0e0f"), dtype=np.uint8)
    pt = np.frombuffer(bytes.fromhex("00112233445566778899aabbccddeeff"), dtype=np.uint8)
    sbox, _ = aes_fhe.tables.sbox_tables()
    c0 = ctx.engine.counters()
    a = pipe.add_round_key(*pipe.encoder.encode(pt), *pipe.encoder.encode(key))
    assert bytes(pipe.encoder.decode(*a)).hex() == "00102030405060708090a0b0c0d0e0f0"    # golden enc.r0.ark
    stride = ctx.engine.slot_count // 16
    assert abs(np.abs(ctx.decrypt(a[0])[::stride][:16]) - 256.0).max() < 1e-3           # H3: XOR output modulus 256
    s = pipe.sub_bytes(*pipe._renorm_pair(*a))
    assert bytes(pipe.encoder.decode(*s)) == bytes(sbox[pt ^ key])
    c1 = ctx.engine.counters()
    if not fused:
        assert c1["mul_cc"] - c0["mul_cc"] == 291             # 291 ct*ct, 162 conjugations
        assert c1["keyswitch"] - c0["keyswitch"] == 291 + 162
    else:
        # XOR4 has odd exponents only: 5 products (2, 3, 4, 5, 7) and 4 conjugations per pruned power base
        # SubBytes baby-step/giant-step: 7 (zeta16 basis) + 1 (b = hi * lift) + 15 babies + 7 giants + 4 bivariate LUTs
        # XOR4 mirror split: two fused LUTs per XOR4, conjugations only on the B side (4) plus one on the mirror half
        assert c1["mul_cc"] - c0["mul_cc"] == 4 * 5 + 2 * 2 + 34   # 4 power bases + 2 XOR4 x 2 LUTs + SubBytes
        assert c1["keyswitch"] - c0["keyswitch"] == 58 + 2 * (4 + 1) + 3   # + conjugations + 3 in SubBytes
    # per-stage slots against the reference-semantics stand-in: stated tolerance 1e-4 on unit-modulus slots
    sctx = aes_fhe.EngineContext(1, mode="cpu", thread_count=1, backend=ss, slot_count=ctx.engine.slot_count)
    sp = make_pipe(sctx)
    ref = sp.sub_bytes(*sp._renorm_pair(*sp.add_round_key(*sp.encoder.encode(pt), *sp.encoder.encode(key))))
    for got, want in zip(s, ref):
        assert np.abs(ctx.decrypt(got)[::stride][:16] - sctx.decrypt(want)[::stride][:16]).max() < 1e-4
    ctx.fused = True


def test_config2_one_round_as_shipped_matches_standin(boot_ctx):
    """configs[1] on the as-shipped flow (column-first ShiftRows + row-major MixColumns): bytes must equal
    the reference-semantics stand-in running the same unchanged flow."""
    which, ctx = boot_ctx
    ctx.fused = False                                           # the reference's calls, one for one
    pipe = make_pipe(ctx)
    rng = np.random.RandomState(0)
    state, key = rng.randint(0, 256, 16).astype(np.uint8), rng.randint(0, 256, 16).astype(np.uint8)
    c0 = ctx.engine.counters()
    out = pipe.encrypt_round(*pipe.encoder.encode(state), *pipe.encoder.encode(key))
    c1 = ctx.engine.counters()
    ctx.fused = True
    sctx = aes_fhe.EngineContext(1, mode="cpu", thread_count=1, backend=ss, slot_count=ctx.engine.slot_count)
    sp = make_pipe(sctx)
    want = sp.encoder.decode(*sp.encrypt_round(*sp.encoder.encode(state), *sp.encoder.encode(key)))
    assert bytes(pipe.encoder.decode(*out)) == bytes(want)
    assert c1["bootstrap"] - c0["bootstrap"] == 2
    assert c1["mul_cc"] - c0["mul_cc"] == 1034 + 2 * 34      # SURVEY App. B round count + EvalMod multiplications


def test_config5_batched_fips_round(boot_ctx):
    """configs[4] shape: every stride position carries an independent block; one FIPS-197 round (R2+R3 driver)."""
    which, ctx = boot_ctx
    pipe = make_pipe(ctx)
    drv = aes_fhe.FipsDriver(pipe, batched=True)
    stride = ctx.engine.slot_count // 16
    rng = np.random.default_rng(5)
    blocks = rng.integers(0, 256, (stride, 16), dtype=np.uint8)
    key = np.frombuffer(bytes.fromhex("000102030405060708090a0b0c0d0e0f"), dtype=np.uint8)
    rks = aes_fhe.expand_aes128_key(key)
    rk_ct = pipe._prepare_round_keys([drv._perm(rk) for rk in rks])
    out = pipe.encrypt_round(*pipe.encoder.encode(drv._perm(blocks)), *rk_ct[1])
    assert np.array_equal(drv.decode(*out), plain_round(blocks, rks[1]))


def test_config5_round_with_both_planes_stacked(boot_ctx, monkeypatch):
    """EngineContext.pair_apply: every step that treats the two nibble planes alike (AddRoundKey, XOR4, renorm, ShiftRows,
    column rotations, bootstrap) runs ONCE on a handle holding both planes (AESFHE_STACK_PAIRS=1; the default stacks the
    bootstrap only).  Same FIPS-197 bytes; round keys (nb = 1) are repeated per item, a plane pair cut from one stacked
    result is re-stacked without a copy."""
    which, ctx = boot_ctx
    import aes_fhe.context as C
    monkeypatch.setattr(C, "STACK_PAIRS", True)
    pipe = make_pipe(ctx)
    drv = aes_fhe.FipsDriver(pipe, batched=True)
    eng = ctx.engine
    stride = eng.slot_count // 16
    rng = np.random.default_rng(77)
    blocks = rng.integers(0, 256, (2, stride, 16), dtype=np.uint8)
    rks = aes_fhe.expand_aes128_key(np.frombuffer(bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c"), dtype=np.uint8))
    rk_ct = pipe._prepare_round_keys([drv._perm(rk) for rk in rks])
    ct = pipe.encoder.encode(drv._perm(blocks))
    k0 = eng.counters()
    out = pipe.encrypt_round(*ct, *rk_ct[1])
    k1 = eng.counters()
    assert out[0].batch == 2 and out[0]._plane_of[0] is out[1]._plane_of[0]          # cut from one stacked handle
    assert k1["bootstrap"] - k0["bootstrap"] == 4                                     # 2 pairs x 2 planes, one call
    for p in range(2):
        assert np.array_equal(drv.decode(*out)[p], plain_round(blocks[p], rks[1]))
    # the default: only the bootstrap is stacked
    monkeypatch.setattr(C, "STACK_PAIRS", False)
    out = pipe.encrypt_round(*ct, *rk_ct[1])
    for p in range(2):
        assert np.array_equal(drv.decode(*out)[p], plain_round(blocks[p], rks[1]))


def test_config5_many_pairs_in_one_batched_handle(boot_ctx):
    """configs[4]: "many ciphertexts" -- P ciphertext pairs travel as ONE batched handle pair (Ciphertext.batch = P), so every
    step of the round (LUTs, key switches, renorms, both bootstraps) runs all pairs through one set of kernel launches;
    the round keys stay unbatched and are broadcast.  Every block of every pair equals the plain FIPS-197 round, eagerly
    and as a captured graph replayed on other inputs."""
    which, ctx = boot_ctx
    pipe = make_pipe(ctx)
    drv = aes_fhe.FipsDriver(pipe, batched=True)
    eng = ctx.engine
    stride = eng.slot_count // 16
    P = 2
    rng = np.random.default_rng(55)
    blocks = rng.integers(0, 256, (P, stride, 16), dtype=np.uint8)
    other = rng.integers(0, 256, (P, stride, 16), dtype=np.uint8)
    key = np.frombuffer(bytes.fromhex("000102030405060708090a0b0c0d0e0f"), dtype=np.uint8)
    rks = aes_fhe.expand_aes128_key(key)
    rk_ct = pipe._prepare_round_keys([drv._perm(rk) for rk in rks])
    ct = pipe.encoder.encode(drv._perm(blocks))
    assert ct[0].batch == P and rk_ct[1][0].batch == 1
    k0 = eng.counters()
    out = pipe.encrypt_round(*ct, *rk_ct[1])
    k1 = eng.counters()
    assert out[0].batch == P and k1["bootstrap"] - k0["bootstrap"] == 2 * P
    got = drv.decode(*out)
    assert got.shape == (P, stride, 16)
    for j in range(P):
        assert np.array_equal(got[j], plain_round(blocks[j], rks[1]))
    rnd = aes_fhe.CapturedRound(pipe, ct, rk_ct[1])
    got = drv.decode(*rnd(*pipe.encoder.encode(drv._perm(other)), *rk_ct[2]))
    for j in range(P):
        assert np.array_equal(got[j], plain_round(other[j], rks[2]))
    rnd.close()


def test_device_zeta16_codec_equals_the_host_codec(boot_ctx):
    """encrypt_zeta16 / decrypt_zeta16 (nibbles over PCIe, codeword lookup and nearest-codeword search on the device)
    against the host codec of the reference (utils.py:9-19): same slots within encryption noise, same nibbles exactly,
    also for slots pushed a third of the way to the neighbouring codeword."""
    which, ctx = boot_ctx
    eng = ctx.engine
    n = eng.slot_count
    rng = np.random.default_rng(4)
    nib = rng.integers(0, 16, n).astype(np.uint8)
    ct = eng.encrypt_zeta16(nib)
    assert ct.level == eng.params()["fresh_level"]
    z = eng.decrypt(ct)
    assert np.abs(z - aes_fhe.to_zeta(nib, 16)).max() < 1e-7
    assert np.array_equal(eng.decrypt_zeta16(ct), nib)
    assert np.array_equal(aes_fhe.from_zeta(z, 16), nib)
    off = aes_fhe.to_zeta(nib, 16) * np.exp(1j * (2 * np.pi / 16) * rng.uniform(-0.33, 0.33, n)) * rng.uniform(0.7, 1.3, n)
    assert np.array_equal(eng.decrypt_zeta16(eng.encrypt(off)), nib)
    assert eng.encrypt_zeta16(nib, level=5).level == 5
    with pytest.raises(ValueError):
        eng.encrypt_zeta16(nib[:-1])


def test_captured_round_replays_on_new_inputs(boot_ctx):
    """The batched FIPS round recorded as one graph (CUDA graph on the device, recorded closures under emulation): replays
    with OTHER blocks and ANOTHER round key must give the plain round of those inputs; the recording run itself executes
    nothing, so a wrong static buffer, a stale memoised level alignment or scratch shared with eager work shows here."""
    which, ctx = boot_ctx
    pipe = make_pipe(ctx)
    drv = aes_fhe.FipsDriver(pipe, batched=True)
    eng = ctx.engine
    stride = eng.slot_count // 16
    rng = np.random.default_rng(21)
    key = np.frombuffer(bytes.fromhex("000102030405060708090a0b0c0d0e0f"), dtype=np.uint8)
    rks = aes_fhe.expand_aes128_key(key)
    rk_ct = pipe._prepare_round_keys([drv._perm(rk) for rk in rks])
    a = rng.integers(0, 256, (stride, 16), dtype=np.uint8)
    b = rng.integers(0, 256, (stride, 16), dtype=np.uint8)
    ct_a = pipe.encoder.encode(drv._perm(a))
    l0 = eng.counters()
    rnd = aes_fhe.CapturedRound(pipe, ct_a, rk_ct[1])
    info = rnd.info()
    assert info["nodes"] > 1000 and info["launches"] > 1000
    k0 = eng.counters()
    out = rnd(*pipe.encoder.encode(drv._perm(b)), *rk_ct[2], stream=1)
    eng.graph_wait(1)
    assert np.array_equal(drv.decode(*out), plain_round(b, rks[2]))
    k1 = eng.counters()
    assert k1["launches"] - k0["launches"] >= info["launches"]          # replays are counted as the launches they are
    assert k1["bootstrap"] - k0["bootstrap"] == 2
    # eager work between replays must not disturb the graph's private arena, nor the other way round
    eager = pipe.add_round_key(*pipe.encoder.encode(drv._perm(a)), *rk_ct[0])
    out = rnd(*ct_a, *rk_ct[1])
    assert np.array_equal(drv.decode(*out), plain_round(a, rks[1]))
    assert np.array_equal(drv.decode(*eager), a ^ rks[0][None, :])
    rnd.close()
    out2 = pipe.encrypt_round(*ct_a, *rk_ct[3])                        # the engine still works eagerly afterwards
    assert np.array_equal(drv.decode(*out2), plain_round(a, rks[3]))


def test_captured_decrypt_round_replays_on_new_inputs(boot_ctx):
    """The README-order decryption round (InvShiftRows, InvSubBytes, AddRoundKey, InvMixColumns with the GF 9/11/13/14 LUTs
    and two bootstraps) recorded as one graph and replayed with other ciphertext bytes and another round key."""
    from bench import plain_inv_round
    from aes_fhe.steps import SHIFTROWS_DEPTH, SUBBYTES_DEPTH
    which, ctx = boot_ctx
    pipe = make_pipe(ctx)
    drv = aes_fhe.FipsDriver(pipe, batched=True)
    stride = ctx.engine.slot_count // 16
    rng = np.random.default_rng(23)
    rks = aes_fhe.expand_aes128_key(np.frombuffer(bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c"), dtype=np.uint8))
    rk_ct = pipe._prepare_round_keys([drv._perm(rk) for rk in rks])
    a = rng.integers(0, 256, (stride, 16), dtype=np.uint8)
    b = rng.integers(0, 256, (stride, 16), dtype=np.uint8)
    lvl = SHIFTROWS_DEPTH + SUBBYTES_DEPTH
    ct_a = pipe.encoder.encode(drv._perm(a), level=lvl)
    rnd = aes_fhe.CapturedRound(pipe, ct_a, rk_ct[9], inverse=True)
    out = rnd(*pipe.encoder.encode(drv._perm(b), level=lvl), *rk_ct[4])
    assert np.array_equal(drv.decode(*out), plain_inv_round(b, rks[4]))
    out = rnd(*ct_a, *rk_ct[9])
    assert np.array_equal(drv.decode(*out), plain_inv_round(a, rks[9]))
    assert out[0].level >= lvl                                   # ready for the next round's InvShiftRows + InvSubBytes
    rnd.close()


@pytest.mark.gpu
def test_config3_config4_with_captured_rounds_gpu():
    """configs[2] and [3] with the nine middle rounds of each direction replayed from ONE recorded round graph
    (`captured=True`): FIPS-197 ciphertext, then the round trip back to the plaintext."""
    mod = backend.use_cuda()
    ctx = aes_fhe.EngineContext(1, mode="gpu", thread_count=1, backend=mod, logn=16, levels=21, fresh_level=14)
    pipe = make_pipe(ctx)
    drv = aes_fhe.FipsDriver(pipe, batched=True)
    stride = ctx.engine.slot_count // 16
    rng = np.random.default_rng(17)
    blocks = rng.integers(0, 256, (stride, 16), dtype=np.uint8)
    blocks[0] = np.frombuffer(bytes.fromhex("00112233445566778899aabbccddeeff"), dtype=np.uint8)
    key = bytes.fromhex("000102030405060708090a0b0c0d0e0f")
    rks = aes_fhe.expand_aes128_key(np.frombuffer(key, dtype=np.uint8))
    l0 = ctx.engine.counters()
    ct = drv.encrypt(blocks, rks, captured=True)
    got = drv.decode(*ct)
    assert bytes(got[0]).hex() == "69c4e0d86a7b0430d8cdb78070b4c55a"                     # FIPS-197 C.1
    want = np.frombuffer(ecb(key, blocks.tobytes()), dtype=np.uint8).reshape(stride, 16)
    assert np.array_equal(got, want)
    assert ctx.engine.counters()["bootstrap"] - l0["bootstrap"] == 2 + 2 + 18            # warm-up, recording, 9 replays
    back = drv.decode(*drv.decrypt(*ct, rks, captured=True))
    assert np.array_equal(back, blocks)
    assert len(pipe._round_graphs) == 6                                                  # first, middle, last round per direction
    pipe.release_graphs()


@pytest.mark.gpu
def test_config3_full_fips_encryption_gpu():
    """configs[2]: full AES-128 encryption of 2048 packed blocks with 18 bootstraps, bit-exact with FIPS-197."""
    mod = backend.use_cuda()
    ctx = aes_fhe.EngineContext(1, mode="gpu", thread_count=1, backend=mod, logn=16, levels=21, fresh_level=14)
    pipe = make_pipe(ctx)
    drv = aes_fhe.FipsDriver(pipe, batched=True)
    stride = ctx.engine.slot_count // 16
    rng = np.random.default_rng(7)
    blocks = rng.integers(0, 256, (stride, 16), dtype=np.uint8)
    blocks[0] = np.frombuffer(bytes.fromhex("00112233445566778899aabbccddeeff"), dtype=np.uint8)
    key = bytes.fromhex("000102030405060708090a0b0c0d0e0f")
    rks = aes_fhe.expand_aes128_key(np.frombuffer(key, dtype=np.uint8))
    got = drv.decode(*drv.encrypt(blocks, rks))
    assert bytes(got[0]).hex() == "69c4e0d86a7b0430d8cdb78070b4c55a"                     # FIPS-197 C.1
    want = np.frombuffer(ecb(key, blocks.tobytes()), dtype=np.uint8).reshape(stride, 16)
    assert np.array_equal(got, want)
    assert ctx.bootstrap_stats()["count"] == 18


@pytest.mark.gpu
def test_config4_roundtrip_decrypt_gpu():
    """configs[3]: encrypt -> decrypt round trip (README-order decryption with InvMixColumns GF9/11/13/14 LUTs and
    InvSubBytes) on 2048 packed blocks: decrypt(FIPS ciphertext) == plaintext, 18 more bootstraps."""
    mod = backend.use_cuda()
    ctx = aes_fhe.EngineContext(1, mode="gpu", thread_count=1, backend=mod, logn=16, levels=21, fresh_level=14)
    pipe = make_pipe(ctx)
    drv = aes_fhe.FipsDriver(pipe, batched=True)
    stride = ctx.engine.slot_count // 16
    rng = np.random.default_rng(11)
    blocks = rng.integers(0, 256, (stride, 16), dtype=np.uint8)
    key = bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c")
    rks = aes_fhe.expand_aes128_key(np.frombuffer(key, dtype=np.uint8))
    cipher = np.frombuffer(ecb(key, blocks.tobytes()), dtype=np.uint8).reshape(stride, 16)
    ct = pipe.encoder.encode(drv._perm(cipher))               # start from the true AES ciphertext
    got = drv.decode(*drv.decrypt(*ct, rks))
    assert np.array_equal(got, blocks)
    assert ctx.bootstrap_stats()["count"] == 18


def test_inverse_round_pieces_on_engine(boot_ctx):
    """InvShiftRows + InvSubBytes + InvMixColumns (fused GF 9/11/13/14 LUTs) on the engine against the plain model."""
    which, ctx = boot_ctx
    pipe = make_pipe(ctx)
    drv = aes_fhe.FipsDriver(pipe, batched=True)
    stride = ctx.engine.slot_count // 16
    rng = np.random.default_rng(3)
    blocks = rng.integers(0, 256, (stride, 16), dtype=np.uint8)
    sbox, isbox = aes_fhe.tables.sbox_tables()
    ct = pipe.encoder.encode(drv._perm(blocks), level=15)
    out = pipe.inv_sub_bytes(*pipe.inv_shift_rows(*ct))
    idx = np.array([(i - 4 * (i % 4)) % 16 for i in range(16)])            # InvShiftRows on column-first bytes
    assert np.array_equal(drv.decode(*out), isbox[blocks[:, idx]])
    mixed = pipe.inv_mix_columns(*pipe.encoder.encode(drv._perm(blocks), level=10))
    gm = lambda m: np.array([aes_fhe.tables.gf_mul(x, m) for x in range(256)], dtype=np.uint8)
    m9, m11, m13, m14 = gm(9), gm(11), gm(13), gm(14)
    want = np.zeros_like(blocks)
    for c in range(4):
        a = [blocks[:, 4 * c + r] for r in range(4)]
        want[:, 4 * c + 0] = m14[a[0]] ^ m11[a[1]] ^ m13[a[2]] ^ m9[a[3]]
        want[:, 4 * c + 1] = m9[a[0]] ^ m14[a[1]] ^ m11[a[2]] ^ m13[a[3]]
        want[:, 4 * c + 2] = m13[a[0]] ^ m9[a[1]] ^ m14[a[2]] ^ m11[a[3]]
        want[:, 4 * c + 3] = m11[a[0]] ^ m13[a[1]] ^ m9[a[2]] ^ m14[a[3]]
    assert np.array_equal(drv.decode(*mixed), want)
