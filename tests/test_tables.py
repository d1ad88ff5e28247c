"""Coefficient tables + key schedule: built from first principles, pinned to FIPS-197 and (when
/root/reference is mounted) to the reference's shipped JSON files gen/coeff/*.json."""
import json

import numpy as np
import pytest

import refload
from aes_fhe import tables


def test_sbox_matches_fips197_known_entries():
    s, inv = tables.sbox_tables()
    # FIPS-197 Fig. 7 corner/known entries
    assert (s[0x00], s[0x01], s[0x53], s[0xFF], s[0x10]) == (0x63, 0x7C, 0xED, 0x16, 0xCA)
    assert np.array_equal(inv[s], np.arange(256))


def test_key_schedule_fips_appendix_a1():
    rks = tables.expand_aes128_key(np.frombuffer(bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c"), dtype=np.uint8))
    assert bytes(rks[1]).hex() == "a0fafe1788542cb123a339392a6c7605"
    assert bytes(rks[10]).hex() == "d014f9a8c9ee2589e13f0cc8b6630ca6"


def test_key_schedule_matches_golden_round_keys():
    from pathlib import Path
    g = json.load(open(Path(__file__).parent / "golden" / "aes_reference_golden.json"))
    for case in g["cases"].values():
        rks = tables.expand_aes128_key(np.frombuffer(bytes.fromhex(case["key"]), dtype=np.uint8))
        assert [bytes(r).hex() for r in rks] == case["round_keys"]


def _resynth2(entries, h, l):
    return sum(c * tables.ZETA16 ** (p * h + q * l) for p, q, c in entries)


@pytest.mark.parametrize("mult", [1, 2, 3, 9, 11, 13, 14])
def test_gf_tables_resynthesise_all_256_inputs(mult):
    # same self-check as gen/generate_gf_mult_2var_coeff.py:80-103
    for which in ("hi", "lo"):
        ent = tables.gf_mult_entries(mult, which)
        for x in range(256):
            z = _resynth2(ent, x >> 4, x & 15)
            y = tables.gf_mul(x, mult)
            want = (y >> 4) if which == "hi" else (y & 15)
            got = int(np.rint(-np.angle(z) * 16 / (2 * np.pi))) % 16
            assert got == want and abs(abs(z) - 1) < 1e-9


def test_xor_table_has_reference_256x_scale_and_64_terms():
    C = tables.xor4_coeffs()
    assert (np.abs(C) > 1e-12).sum() == 64
    assert abs(np.abs(C).max() - 95.47) < 0.01          # 256 x the unit-modulus table
    a, b = 5, 12
    z = sum(C[p, q] * tables.ZETA16 ** (p * a + q * b) for p in range(16) for q in range(16))
    assert abs(z - 256 * tables.ZETA16 ** (a ^ b)) < 1e-9


def test_sbox_poly_evaluates_to_sbox():
    s, inv = tables.sbox_tables()
    w = np.exp(-2j * np.pi / 256)
    for inverse, tab in ((False, s), (True, inv)):
        hi, lo = tables.sbox_coeffs(inverse)
        x = np.arange(256)
        V = w ** (np.outer(x, np.arange(len(hi))) % 256)
        zh, zl = V @ hi, V @ lo
        nib = lambda z: (np.rint(-np.angle(z) * 16 / (2 * np.pi)).astype(int)) % 16
        assert np.array_equal(nib(zh) * 16 + nib(zl), tab)


@pytest.mark.skipif(not refload.available(), reason="reference tree not mounted")
def test_tables_equal_reference_json():
    d = refload.REF / "gen" / "coeff"
    ref = json.load(open(d / "xor4_coeffs.json"))["entries"]
    C = tables.xor4_coeffs()
    assert [(p, q) for p, q, *_ in ref] == [(p, q) for p in range(16) for q in range(16) if abs(C[p, q]) > 1e-12]
    assert max(abs(C[p, q] - complex(re, im)) for p, q, re, im in ref) < 1e-12
    for name, (inverse, idx) in {"mod256_to_16_hi": (False, 0), "mod256_to_16_lo": (False, 1),
                                "inv_mod256_to_16_hi": (True, 0), "inv_mod256_to_16_lo": (True, 1)}.items():
        ent = json.load(open(d / f"{name}.json"))["entries"]
        mine = tables.sbox_coeffs(inverse)[idx]
        assert [k for k, *_ in ent] == [k for k, c in enumerate(mine) if abs(c) > 1e-12]
        assert max(abs(mine[k] - complex(re, im)) for k, re, im in ent) < 1e-13
    for mult in (1, 2, 3, 9, 11, 13, 14):
        for which in ("hi", "lo"):
            ent = json.load(open(d / f"gf_mult{mult}_{which}_coeffs.json"))["entries"]
            mine = tables.gf_mult_entries(mult, which)
            assert [(p, q) for p, q, *_ in ent] == [(p, q) for p, q, _ in mine]
            assert max(abs(c - complex(re, im)) for (p, q, re, im), (_, _, c) in zip(ent, mine)) < 1e-13
