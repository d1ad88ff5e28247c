"""LUT polynomial coefficient tables for the Zeta16 nibble encoding (ePrint 2024/274).

Built from first principles (GF(2^8) arithmetic + inverse DFTs of `zeta16**f(x)`),
not read from the reference's JSON files.  The values reproduce what the reference
ships under `gen/coeff/*.json` (generators: `gen/generate_xor4_coeffs.py:13-45`,
`gen/generate_sobx_coeffs.py:66-78`, `gen/generate_gf_mult_2var_coeff.py:31-60`),
including the reference's 256x scale on the XOR table and the
row-major (p outer, q inner) ordering of the sparse entries, which fixes the order of
engine calls in the LUT evaluators.  `tests/test_tables.py` checks them against the
reference files when `/root/reference` is present.
"""
from __future__ import annotations

from functools import lru_cache
from typing import Dict, List, Tuple

import numpy as np

ZETA16 = np.exp(-2j * np.pi / 16)


# ---------------------------------------------------------------- GF(2^8) / AES tables
def gf_mul(a: int, b: int) -> int:
    """Carry-less multiply modulo x^8+x^4+x^3+x+1."""
    acc = 0
    for _ in range(8):
        if b & 1:
            acc ^= a
        a = ((a << 1) ^ (0x11B if a & 0x80 else 0)) & 0x1FF
        a &= 0xFF
        b >>= 1
    return acc


@lru_cache(maxsize=None)
def sbox_tables() -> Tuple[np.ndarray, np.ndarray]:
    """FIPS-197 S-box and its inverse, computed (inverse in GF(2^8) + affine map)."""
    inv = [0] * 256
    for x in range(1, 256):
        # x^254 = x^-1
        y, e, base = 1, 254, x
        while e:
            if e & 1:
                y = gf_mul(y, base)
            base = gf_mul(base, base)
            e >>= 1
        inv[x] = y
    sbox = np.zeros(256, dtype=np.uint8)
    for x in range(256):
        b = inv[x]
        r = b
        for s in (1, 2, 3, 4):
            r ^= ((b << s) | (b >> (8 - s))) & 0xFF
        sbox[x] = r ^ 0x63
    isbox = np.zeros(256, dtype=np.uint8)
    isbox[sbox] = np.arange(256, dtype=np.uint8)
    return sbox, isbox


RCON = np.array([0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1B, 0x36], dtype=np.uint8)


def expand_aes128_key(master: np.ndarray) -> List[np.ndarray]:
    """AES-128 key schedule -> 11 round keys of 16 bytes, FIPS byte order
    (same outputs as the reference test driver, `test/test_aes_pipeline_roundtrip.py:95-110`)."""
    master = np.asarray(master, dtype=np.uint8)
    assert master.shape == (16,)
    sbox, _ = sbox_tables()
    words = [master[4 * i:4 * i + 4].copy() for i in range(4)]
    for i in range(4, 44):
        t = words[i - 1].copy()
        if i % 4 == 0:
            t = sbox[np.roll(t, -1)]
            t[0] ^= RCON[i // 4 - 1]
        words.append(words[i - 4] ^ t)
    return [np.concatenate(words[4 * r:4 * r + 4]).astype(np.uint8) for r in range(11)]


# ---------------------------------------------------------------- coefficient tables
@lru_cache(maxsize=None)
def xor4_coeffs(normalized: bool = False) -> np.ndarray:
    """16x16 complex table c[p,q] with sum_pq c[p,q] z^(p a) z^(q b) = 256 * z^(a xor b).

    The factor 256 is the reference's (`gen/generate_xor4_coeffs.py:13-17` multiplies
    ifft2 by n^2); entries below 1e-8 are dropped there and below 1e-12 by the consumer.

    `normalized=True` is the corrected table: the plain inverse DFT, whose outputs are unit-modulus
    codewords z^(a xor b).  With it an XOR output can feed the next LUT (or a bootstrap) directly, without the
    decrypt / re-encrypt "hard renorm" the reference needs after every XOR (pipeline.py:65-69) -- the precondition of
    the keyless flow built on `aes_fhe.snap` (tests/test_snap.py::test_normalized_xor_chain_needs_no_renorm)."""
    a = np.arange(16)
    F = ZETA16 ** (a[:, None] ^ a[None, :])
    C = np.fft.ifft2(F) * (1.0 if normalized else 256.0)
    C[np.abs(C) <= (1e-8 / 256.0 if normalized else 1e-8)] = 0
    return C


def _lut1d(table: np.ndarray) -> np.ndarray:
    samples = ZETA16 ** table.astype(np.int64)
    c = np.fft.ifft(samples.astype(np.complex128))
    keep = np.abs(c) > 1e-12
    kmax = int(np.nonzero(keep)[0].max())
    out = np.where(keep, c, 0)[:kmax + 1]
    return out


@lru_cache(maxsize=None)
def sbox_coeffs(inverse: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """(hi, lo) degree-255 coefficient vectors: sum_k c_k w^(k x) = zeta16^nibble(S[x]), w = zeta256."""
    sbox, isbox = sbox_tables()
    t = isbox if inverse else sbox
    return _lut1d((t >> 4) & 0xF), _lut1d(t & 0xF)


@lru_cache(maxsize=None)
def gf_mult_entries(mult: int, which: str) -> Tuple[Tuple[int, int, complex], ...]:
    """Sparse bivariate table for y = mult * x in GF(2^8): entries (p, q, c) with
    sum c z^(p h) z^(q l) = z^nibble_which(y), x = 16 h + l; row-major order, |c| > 1e-12."""
    if which not in ("hi", "lo"):
        raise ValueError("which must be 'hi' or 'lo'")
    S = np.empty((16, 16), dtype=np.complex128)
    for h in range(16):
        for l in range(16):
            y = gf_mul((h << 4) | l, mult)
            S[h, l] = ZETA16 ** ((y >> 4) & 0xF if which == "hi" else y & 0xF)
    C = np.fft.ifft2(S)
    return tuple((p, q, complex(C[p, q])) for p in range(16) for q in range(16) if abs(C[p, q]) > 1e-12)


def load_all_coeffs() -> Dict[str, np.ndarray]:
    """Dictionary with the keys `AESPipeline` expects (`pipeline.py:19-26`)."""
    hi, lo = sbox_coeffs(False)
    ihi, ilo = sbox_coeffs(True)
    return {"xor4": xor4_coeffs(), "sub_hi": hi, "sub_lo": lo, "inv_sub_hi": ihi, "inv_sub_lo": ilo}
