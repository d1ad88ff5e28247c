"""AES-128 round steps on Zeta16-encoded CKKS ciphertext pairs (hi nibble, lo nibble).

Host-side mirror of the reference's L2 step classes.  Each class keeps the reference's
name, constructor and call signature and issues the *same sequence of engine calls*
through `EngineContext` (verified op-for-op by `tests/test_aes_mirror.py` against
the reference's own files on a tracing backend), so that the workload measured on the
B200 engine is the reference's workload:

  StateEncoder        <- state_encoder.py:8-38     XOR4LUT          <- xor4_lut.py:10-77
  AddRoundKey         <- add_round_key.py:138-144  SubBytesLUT      <- sub_bytes_lut.py:8-73
  ShiftRows           <- shift_rows.py:7-56        InvShiftRows     <- inv_shiftrows.py:9-47
  MixColFinal         <- mixcol_final.py:41-165    InvMixColumnsFHE <- invmixcolumns_fhe.py:34-170
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import tables
from .context import EngineContext

Pair = Tuple[Any, Any]

import os as _os
# tuning switches for the stream-lane granularity (measured choices recorded in profiles/README.md)
LANES_CONJ = _os.environ.get("AESFHE_LANES_CONJ", "1") == "1"
LANES_AB = _os.environ.get("AESFHE_LANES_AB", "1") == "1"
SHARE_GF_BASES = _os.environ.get("AESFHE_SHARE_GF_BASES", "1") == "1"   # fused mode: M_k(rot_k x) = rot_k(M_k x)
XOR_TREE = _os.environ.get("AESFHE_XOR_TREE", "0") == "1"       # fused MixColumns: (a ^ b) ^ (c ^ d) on two lanes; measured: no gain
XOR_MIRROR = _os.environ.get("AESFHE_XOR_MIRROR", "1") == "1"   # fused XOR4: no conjugations on the A side
PRUNE_BASIS = _os.environ.get("AESFHE_PRUNE_BASIS", "1") == "1"   # fused mode: build only the powers a LUT uses

# multiplicative depths of the steps, used for the encryption-level hints of the fused mode
XOR4_DEPTH = 5          # power basis 3 + product 1 + constant 1
GF_DEPTH = 5
# fused mode evaluates the two degree-255 S-box polynomials baby-step/giant-step (b^k = b^i (b^16)^j, one fused bivariate
# LUT per half: 39 key switches instead of 138) at the price of one more level than the reference's 13 (3+1+1+7+1)
PS_SUBBYTES = _os.environ.get("AESFHE_PS_SUBBYTES", "1") == "1"
SUBBYTES_DEPTH = 14 if PS_SUBBYTES else 13
SHIFTROWS_DEPTH = 1


def to_zeta(values: np.ndarray, modulus: int = 16) -> np.ndarray:
    """k -> exp(-2 pi i k / modulus)   (utils.py:9-12)."""
    return np.exp(-2j * np.pi / modulus) ** (np.asarray(values) % modulus)


def from_zeta(z: np.ndarray, modulus: int = 16) -> np.ndarray:
    """Nearest codeword index from the argument only (utils.py:15-19)."""
    k = -np.angle(z) * modulus / (2 * np.pi)
    return np.mod(np.rint(k), modulus).astype(np.uint8)


class StateEncoder:
    """16 bytes <-> two ciphertexts; byte i sits in slot i*stride, every other slot is 1.0."""

    def __init__(self, ctx: EngineContext):
        self.ctx = ctx
        self.sc = ctx.engine.slot_count
        self.stride = self.sc // 16

    def encode(self, state: np.ndarray, level=None) -> Pair:
        assert state.shape == (16,)
        pos = np.arange(16) * self.stride
        vecs = []
        for nib in ((state >> 4) & 0x0F, state & 0x0F):
            v = np.ones(self.sc, dtype=np.complex128)
            v[pos] = to_zeta(nib.astype(np.uint8), 16)
            vecs.append(v)
        if level is None:
            return self.ctx.encrypt(vecs[0]), self.ctx.encrypt(vecs[1])
        return self.ctx.encrypt(vecs[0], level=level), self.ctx.encrypt(vecs[1], level=level)

    def renorm(self, ct_hi, ct_lo, level=None) -> Pair:
        """encode(decode(hi, lo)): on the B200 engine without leaving the device (slots off the stride grid -> 1.0)."""
        if getattr(self.ctx, "device_renorm", False):
            return self.ctx.pair_apply(self.ctx.snap_zeta16, (ct_hi, level, self.stride), (ct_lo, level, self.stride))
        return self.encode(self.decode(ct_hi, ct_lo), level=level)

    def decode(self, ct_hi, ct_lo) -> np.ndarray:
        pos = np.arange(16) * self.stride
        hi = from_zeta(self.ctx.decrypt(ct_hi)[pos], 16)
        lo = from_zeta(self.ctx.decrypt(ct_lo)[pos], 16)
        return ((hi.astype(np.uint8) << 4) | lo).astype(np.uint8)


def _const_pt(ctx: EngineContext, sc: int, c: complex):
    return ctx.encode(np.full(sc, c, dtype=np.complex128))


def _zeta16_basis_pruned(eng, ct, exps) -> Dict[int, Any]:
    """Fused mode: the zeta16 basis {k: ct^k} restricted to the exponents a LUT really uses.  Powers 9..15 are
    conjugates of 7..1 (unit-modulus slots); only the needed positive powers and only the needed conjugations are
    computed -- XOR4 has odd exponents only: 5 products + 4 conjugations per base instead of 7 + 7."""
    exps = sorted(set(int(k) for k in exps))
    pos_need = sorted({k for k in exps if 1 <= k <= 8} | {16 - k for k in exps if k > 8})
    pos = eng.make_power_basis_sparse(ct, 8, pos_need) if pos_need else []
    basis: Dict[int, Any] = {}
    if 0 in exps:
        basis[0] = eng.add_plain(eng.multiply(ct, 0.0), 1.0)
    basis.update({k: pos[k - 1] for k in exps if 1 <= k <= 8})
    hi = [k for k in exps if k > 8]
    if hi:
        # conjugate at the level the LUT will run at (the deepest needed power): the alignment is memoised on the
        # positive power, so the LUT reuses it, the conjugation key-switches fewer limbs, and conj(x^k) needs no
        # alignment of its own
        deep = min(pos[k - 1].level for k in pos_need)
        src = [eng.level_down(pos[15 - k], deep) for k in hi]
        if LANES_CONJ:
            conj = eng.lane_map(eng.conjugate, [(c,) for c in src])
        else:
            conj = [eng.conjugate(c) for c in src]
        basis.update(dict(zip(hi, conj)))
    return basis


class XOR4LUT:
    """a xor b on nibbles as sum_{p,q} c[p,q] A^p B^q over the zeta16 power bases."""

    def __init__(self, ctx: EngineContext, coeffs: np.ndarray):
        self.ctx = ctx
        self.sc = ctx.engine.slot_count
        self.coeffs = coeffs
        self.pt: Dict[Tuple[int, int], Any] = {
            (p, q): _const_pt(ctx, self.sc, coeffs[p, q])
            for p in range(16) for q in range(16) if abs(coeffs[p, q]) > 1e-12
        }
        self.terms = [(p, q, complex(coeffs[p, q])) for (p, q) in self.pt]
        self.exps_a = sorted({p for p, _, _ in self.terms})
        self.exps_b = sorted({q for _, q, _ in self.terms})
        # mirror split (fused mode): terms with p <= 8 as they are, terms with p > 8 conjugated onto A^(16-p) B^(16-q)
        self._terms_direct = [(p, q, c) for p, q, c in self.terms if 1 <= p <= 8]
        self._terms_mirror = [(16 - p, 16 - q, complex(np.conj(c))) for p, q, c in self.terms if p > 8]
        self._exps_a_pos = sorted({p for p, _, _ in self._terms_direct} | {p for p, _, _ in self._terms_mirror})
        need_b = {q for _, q, _ in self._terms_direct} | {q for _, q, _ in self._terms_mirror}
        self._mirror_ok = (bool(self._terms_direct) and bool(self._terms_mirror) and need_b <= set(self.exps_b)
                           and all(1 <= q <= 15 for q in need_b) and all(p != 0 for p, _, _ in self.terms))

    def _build_power_basis_16(self, ct, exps=None) -> Dict[int, Any]:
        eng = self.ctx
        if exps is not None and getattr(eng, "fused", False) and PRUNE_BASIS:
            return _zeta16_basis_pruned(eng, ct, exps)

        def quiet_intt(x):
            try:
                return eng.to_intt(x)
            except RuntimeError:
                return x

        # recovery ladder of xor4_lut.py:31-51: plain try, coefficient form, bootstrap
        try:
            pos = eng.make_power_basis(ct, 8)
        except RuntimeError:
            ct = quiet_intt(ct)
            try:
                pos = eng.make_power_basis(ct, 8)
            except RuntimeError:
                ct = eng.bootstrap(quiet_intt(ct))
                pos = eng.make_power_basis(ct, 8)
        basis = {0: eng.add_plain(eng.sub(ct, ct), 1.0)}   # "encrypted 1" without spending a level
        basis.update({k: pos[k - 1] for k in range(1, 9)})
        if getattr(eng, "fused", False) and LANES_CONJ:      # the seven conjugations are independent
            conj = eng.lane_map(eng.conjugate, [(pos[15 - k],) for k in range(9, 16)])
            basis.update({k: conj[k - 9] for k in range(9, 16)})
            return basis
        for k in range(9, 16):
            basis[k] = eng.conjugate(pos[15 - k])
        return basis

    def apply(self, a_ct, b_ct):
        eng = self.ctx
        if getattr(eng, "fused", False):
            # build both power bases at the common level (a cached round-key ciphertext may sit far above the state)
            lvl = min(a_ct.level, b_ct.level)
            a_ct, b_ct = eng.level_down(a_ct, lvl), eng.level_down(b_ct, lvl)
        if getattr(eng, "fused", False) and XOR_MIRROR and PRUNE_BASIS and LANES_AB and self._mirror_ok:
            # A^p for p > 8 is conj(A^(16-p)):  sum_{p>8} c A^p B^q = conj( sum conj(c) A^(16-p) B^(16-q) ), so only the
            # positive powers of A are needed (no conjugations on the A side): two fused LUTs over the same bases and ONE
            # conjugation instead of four
            A, B = eng.pair_map(self._build_power_basis_16, (a_ct, self._exps_a_pos), (b_ct, self.exps_b))
            la, lb = [A.get(k) for k in range(16)], [B.get(k) for k in range(16)]
            direct, mirror = eng.pair_map(eng.lut2, (la, lb, self._terms_direct), (la, lb, self._terms_mirror))
            return eng.add(direct, eng.conjugate(mirror))
        if getattr(eng, "fused", False) and LANES_AB:
            A, B = eng.pair_map(self._build_power_basis_16, (a_ct, self.exps_a), (b_ct, self.exps_b))   # two stream lanes
        else:
            fused = getattr(eng, "fused", False)
            A = self._build_power_basis_16(a_ct, self.exps_a if fused else None)
            B = self._build_power_basis_16(b_ct, self.exps_b if fused else None)
        if getattr(eng, "fused", False):
            # same polynomial, one tensor accumulation and ONE relinearisation (csrc/lut.cu) instead of 64
            return eng.lut2([A.get(k) for k in range(16)], [B.get(k) for k in range(16)], self.terms)
        acc = eng.sub(A[0], A[0])
        for (p, q), pt in self.pt.items():
            prod = eng.multiply(A[p], B[q])
            acc = eng.add(acc, eng.multiply(prod, pt))
        return acc

    __call__ = apply


class AddRoundKey:
    def __init__(self, xor4: XOR4LUT):
        self.xor4 = xor4

    def __call__(self, ct_hi, ct_lo, key_hi, key_lo) -> Pair:
        ctx = self.xor4.ctx
        if getattr(ctx, "fused", False):      # the two nibble planes never interact inside XOR4: overlap them
            return ctx.pair_apply(self.xor4.apply, (ct_hi, key_hi), (ct_lo, key_lo))
        return self.xor4.apply(ct_hi, key_hi), self.xor4.apply(ct_lo, key_lo)


class SubBytesLUT:
    """S-box as two degree-255 polynomials in b = zeta256^byte sharing one power basis.
    (reference class `SubBytesLUTFastCached`, imported by pipeline.py:9 as `SubBytesLUT`)."""

    def __init__(self, ctx: EngineContext, hi_coeffs: np.ndarray, lo_coeffs: np.ndarray):
        self.ctx = ctx
        self.sc = ctx.engine.slot_count
        self.hi = np.array(hi_coeffs, dtype=np.complex128)
        self.lo = np.array(lo_coeffs, dtype=np.complex128)
        tol = 1e-12
        self.ks_hi = [k for k, c in enumerate(self.hi) if abs(c) > tol]
        self.ks_lo = [k for k, c in enumerate(self.lo) if abs(c) > tol]
        union = sorted(set(self.ks_hi) | set(self.ks_lo))
        self.ks_union = [k for k in union if k != 0]
        self.deg256 = min(max(union) if union else 0, 128)
        self.pt_hi = {k: _const_pt(ctx, self.sc, self.hi[k]) for k in self.ks_hi}
        self.pt_lo = {k: _const_pt(ctx, self.sc, self.lo[k]) for k in self.ks_lo}
        self.c0_hi = self.hi[0] if len(self.hi) else 0j
        self.c0_lo = self.lo[0] if len(self.lo) else 0j
        # lift polynomial zeta16^l -> zeta256^l (16-point inverse DFT)
        w256 = np.exp(-2j * np.pi / 256)
        lift = np.fft.ifft(np.array([w256 ** k for k in range(16)], dtype=np.complex128))
        self.ks_lift = [k for k, c in enumerate(lift) if abs(c) > tol and k != 0]
        self.deg16 = min(max(self.ks_lift) if self.ks_lift else 0, 8)
        self.pt_lift = {k: _const_pt(ctx, self.sc, lift[k]) for k in self.ks_lift}
        self.c0_lift = lift[0]
        self._lift = lift
        self.ks_hi_nz = [k for k in self.ks_hi if k != 0]
        self.ks_lo_nz = [k for k in self.ks_lo if k != 0]
        # baby-step/giant-step term lists (fused mode): b^k = b^i G^j, k = 16 j + i; powers above 128 are conjugates of
        # b^(256-k), so each polynomial is a "direct" bivariate LUT plus the conjugate of a "mirror" one
        def split(coeffs, ks):
            direct = [(k % 16, k // 16, complex(coeffs[k])) for k in ks if 1 <= k <= 128]
            mirror = [((256 - k) % 16, (256 - k) // 16, complex(np.conj(coeffs[k]))) for k in ks if k > 128]
            return direct, mirror
        self._ps_hi = split(self.hi, self.ks_hi_nz)
        self._ps_lo = split(self.lo, self.ks_lo_nz)

    def _poly_fused(self, pos: List[Any], period: int, coeffs, ks, c0):
        """c0 + sum_k c_k X^k with X^(period-k) = conj(X^k):  sum_{k<=len(pos)} c_k X^k + conj(sum conj(c_k) X^(period-k)).
        One fused linear combination per half and ONE conjugation, instead of one conjugation per high power."""
        eng = self.ctx
        direct = [k for k in ks if k <= len(pos)]
        mirror = [k for k in ks if k > len(pos)]
        acc = eng.lincomb([pos[k - 1] for k in direct], [coeffs[k] for k in direct]) if direct else None
        if mirror:
            m = eng.conjugate(eng.lincomb([pos[period - k - 1] for k in mirror], [np.conj(coeffs[k]) for k in mirror]))
            acc = m if acc is None else eng.add(acc, m)
        return eng.add_plain(acc, c0)

    def _apply_bsgs(self, ct_b) -> Pair:
        """Both S-box polynomials from 15 baby powers b^1..b^16 and 7 giant powers (b^16)^1..8: every monomial
        c_k b^(16j+i) is a term (i, j, c_k) of a bivariate LUT over the two bases, evaluated by the engine's fused
        `lut2` with ONE relinearisation -- 22 products + 4 LUTs + 2 conjugations instead of 127 products."""
        eng = self.ctx
        babies = eng.make_power_basis(ct_b, 16)
        giants = eng.make_power_basis(babies[15], 8)
        one = eng.add_plain(eng.multiply(ct_b, 0.0), 1.0)
        A = [one] + list(babies[:15])
        B = [one] + list(giants) + [None] * 7
        jobs = [t for t in (self._ps_hi[0], self._ps_hi[1], self._ps_lo[0], self._ps_lo[1])]
        outs = eng.lane_map(lambda terms: eng.lut2(A, B, terms) if terms else None, [(t,) for t in jobs])
        mir = [o for o in (outs[1], outs[3]) if o is not None]
        conj = iter(eng.lane_map(eng.conjugate, [(o,) for o in mir]))
        res = []
        for direct, mirror, c0 in ((outs[0], outs[1], self.c0_hi), (outs[2], outs[3], self.c0_lo)):
            m = next(conj) if mirror is not None else None
            acc = direct if m is None else (m if direct is None else eng.add(direct, m))
            res.append(eng.add_plain(acc, c0))
        return res[0], res[1]

    def apply(self, ct_hi, ct_lo) -> Pair:
        eng = self.ctx
        if getattr(eng, "fused", False):
            pos16 = eng.make_power_basis(ct_lo, self.deg16)
            lift = self._lift
            lifted = self._poly_fused(pos16, 16, lift, self.ks_lift, self.c0_lift)
            ct_b = eng.multiply(ct_hi, lifted)
            if PS_SUBBYTES and self.deg256 == 128:
                return self._apply_bsgs(ct_b)
            pos256 = eng.make_power_basis(ct_b, self.deg256)
            return eng.pair_map(self._poly_fused, (pos256, 256, self.hi, self.ks_hi_nz, self.c0_hi),
                                (pos256, 256, self.lo, self.ks_lo_nz, self.c0_lo))
        lifted = eng.add_plain(eng.multiply(ct_lo, 0.0), self.c0_lift)
        pos16 = eng.make_power_basis(ct_lo, self.deg16) if self.deg16 > 0 else []
        for k in self.ks_lift:
            bk = pos16[k - 1] if k <= len(pos16) else eng.conjugate(pos16[15 - k])
            lifted = eng.add(lifted, eng.multiply(bk, self.pt_lift[k]))
        ct_b = eng.multiply(ct_hi, lifted)                       # zeta256^byte
        pos256 = eng.make_power_basis(ct_b, self.deg256) if self.deg256 > 0 else []
        out_hi = eng.add_plain(eng.multiply(ct_b, 0.0), self.c0_hi)
        out_lo = eng.add_plain(eng.multiply(ct_b, 0.0), self.c0_lo)
        for k in self.ks_union:
            bk = pos256[k - 1] if k <= len(pos256) else eng.conjugate(pos256[255 - k])
            if k in self.pt_hi:
                out_hi = eng.add(out_hi, eng.multiply(bk, self.pt_hi[k]))
            if k in self.pt_lo:
                out_lo = eng.add(out_lo, eng.multiply(bk, self.pt_lo[k]))
        return out_hi, out_lo


class _MaskedRowRotate:
    """sum_r rotate(ct * rowmask_r, sign * 4 r stride) for column-first packing (row r = slots r+4c)."""

    sign = -1

    def __init__(self, ctx: EngineContext):
        self.ctx = ctx
        self.sc = ctx.engine.slot_count
        self.stride = self.sc // 16
        self._pt_masks: List[Any] = []
        for r in range(4):
            m = np.zeros(self.sc, dtype=np.complex128)
            m[(r + 4 * np.arange(4)) * self.stride] = 1.0
            self._pt_masks.append(ctx.encode(m))
        self._rot_steps = [self.sign * r * 4 * self.stride for r in range(4)]

    def _apply_one(self, ct):
        eng = self.ctx
        out = eng.multiply(ct, 0.0)
        for mask, step in zip(self._pt_masks, self._rot_steps):
            part = eng.multiply(ct, mask)
            if step != 0:
                part = eng.rotate(part, step)
            out = eng.add(out, part)
        return out

    def apply(self, ct_hi, ct_lo) -> Pair:
        if getattr(self.ctx, "fused", False):
            return self.ctx.pair_apply(self._apply_one, (ct_hi,), (ct_lo,))
        return self._apply_one(ct_hi), self._apply_one(ct_lo)


class ShiftRows(_MaskedRowRotate):
    sign = -1


class InvShiftRows(_MaskedRowRotate):
    sign = +1


class _GFTables:
    """Lazy (mult, which) -> {(p,q): plaintext} cache (mixcol_final.py:19-37)."""

    def __init__(self):
        self.pt_cache: Dict[Tuple[int, str], Dict[Tuple[int, int], Any]] = {}

    def load_plaintexts(self, ctx: EngineContext, mult: int, which: str):
        key = (mult, which)
        if key not in self.pt_cache:
            sc = ctx.engine.slot_count
            self.pt_cache[key] = {(p, q): _const_pt(ctx, sc, c) for p, q, c in tables.gf_mult_entries(mult, which)}
        return self.pt_cache[key]


class _MixBase:
    """Shared machinery of MixColFinal / InvMixColumnsFHE: zeta16 bases, bivariate GF LUTs,
    row-major column shifts, XOR accumulation with hard renorm, final bootstraps."""

    def __init__(self, ctx: EngineContext, xor4: XOR4LUT):
        self.ctx = ctx
        self.xor4 = xor4
        self.sc = ctx.engine.slot_count
        self._coeffs = _GFTables()

    def _basis16(self, ct, exps=None) -> Dict[int, Any]:
        eng = self.ctx
        if exps is not None and getattr(eng, "fused", False) and PRUNE_BASIS:
            return _zeta16_basis_pruned(eng, ct, exps)
        try:
            pos = eng.make_power_basis(ct, 8)
        except RuntimeError:
            ct = eng.bootstrap(ct)
            pos = eng.make_power_basis(ct, 8)
        basis = {0: eng.add_plain(eng.multiply(ct, 0.0), 1.0)}
        basis.update({k: pos[k - 1] for k in range(1, 9)})
        if getattr(eng, "fused", False) and LANES_CONJ:
            conj = eng.lane_map(eng.conjugate, [(pos[15 - k],) for k in range(9, 16)])
            basis.update({k: conj[k - 9] for k in range(9, 16)})
            return basis
        for k in range(9, 16):
            basis[k] = eng.conjugate(pos[15 - k])
        return basis

    def _eval2(self, ct_hi, ct_lo, mult: int, which: str, bases=None):
        eng = self.ctx
        if getattr(eng, "fused", False):
            bx, by = bases if bases is not None else (self._basis16(ct_hi), self._basis16(ct_lo))
            terms = [(p, q, complex(c)) for p, q, c in tables.gf_mult_entries(mult, which)]
            return eng.lut2([bx.get(k) for k in range(16)], [by.get(k) for k in range(16)], terms)
        bx = self._basis16(ct_hi)
        by = self._basis16(ct_lo)
        pts = self._coeffs.load_plaintexts(eng, mult, which)
        acc = eng.multiply(ct_hi, 0.0)
        for (p, q), pt in pts.items():
            t = eng.multiply(bx[p], by[q])
            t = eng.multiply(t, pt)
            acc = eng.add(acc, t)
        return acc

    def _gf(self, mult: int, ct_hi, ct_lo) -> Pair:
        if getattr(self.ctx, "fused", False):
            # the reference rebuilds both 16-power bases for the hi and the lo table (mixcol_final.py:82-83);
            # they are identical, so the fused path builds them once
            ents = [e for which in ("hi", "lo") for e in tables.gf_mult_entries(mult, which)]
            bases = self.ctx.pair_map(self._basis16, (ct_hi, {p for p, _, _ in ents}), (ct_lo, {q for _, q, _ in ents}))
            return self.ctx.pair_map(self._eval2, (ct_hi, ct_lo, mult, "hi", bases), (ct_hi, ct_lo, mult, "lo", bases))
        return self._eval2(ct_hi, ct_lo, mult, "hi"), self._eval2(ct_hi, ct_lo, mult, "lo")

    def _gf_shared(self, mults, ct_hi, ct_lo) -> List[Pair]:
        """Fused mode: several GF(2^8) constant multiplications of the SAME state from ONE pair of zeta16 bases.
        The LUTs act slot by slot, so M(rot_k(x)) = rot_k(M(x)): the reference's M_k(rot_k(x)) (mixcol_final.py:124-131,
        invmixcolumns_fhe.py:140-147) is evaluated as rot_k(M_k(x)) -- one basis pair per MixColumns instead of one per
        multiplier, and the rotations act on the (lower-level) LUT outputs."""
        ents = [e for m in mults for which in ("hi", "lo") for e in tables.gf_mult_entries(m, which)]
        bases = self.ctx.pair_map(self._basis16, (ct_hi, {p for p, _, _ in ents}), (ct_lo, {q for _, q, _ in ents}))
        jobs = [(ct_hi, ct_lo, m, which, bases) for m in mults for which in ("hi", "lo")]
        outs = self.ctx.lane_map(self._eval2, jobs)
        return [(outs[2 * i], outs[2 * i + 1]) for i in range(len(mults))]

    def _rot_pair(self, pair: Pair, k_up: int) -> Pair:
        step = -4 * k_up * self.stride
        return self.ctx.pair_apply(self.ctx.rotate, (pair[0], step), (pair[1], step))

    def _col_shift_rowmajor(self, ct, k_up: int):
        return self.ctx.rotate(ct, -4 * k_up * self.stride)

    def _shifts(self, ct_hi, ct_lo):
        if getattr(self.ctx, "fused", False):      # rot1..3 of one ciphertext share one ModUp (hoisting)
            steps = [-4 * k * self.stride for k in (1, 2, 3)]
            return list(zip(*self.ctx.pair_apply(self.ctx.rotate_many, (ct_hi, steps), (ct_lo, steps))))
        return [(self._col_shift_rowmajor(ct_hi, k), self._col_shift_rowmajor(ct_lo, k)) for k in (1, 2, 3)]

    def _xor_pair(self, a: Pair, b: Pair) -> Pair:
        if getattr(self.ctx, "fused", False):
            return self.ctx.pair_apply(self.xor4.apply, (a[0], b[0]), (a[1], b[1]))
        return self.xor4.apply(a[0], b[0]), self.xor4.apply(a[1], b[1])


class MixColFinal(_MixBase):
    """out = 2*x xor 3*rot1(x) xor rot2(x) xor rot3(x), rot_k = rotate by -4k*stride."""

    def __init__(self, ctx: EngineContext, xor4: XOR4LUT, stride: Optional[int] = None):
        super().__init__(ctx, xor4)
        self.stride = stride if stride is not None else self.sc // 16
        self.enc = StateEncoder(ctx)
        self.zero_hi, self.zero_lo = self.enc.encode(np.zeros(16, dtype=np.uint8))

    def gf_mult_2(self, ct_hi, ct_lo) -> Pair:
        return self._gf(2, ct_hi, ct_lo)

    def gf_mult_3(self, ct_hi, ct_lo) -> Pair:
        return self._gf(3, ct_hi, ct_lo)

    def _renorm_pair(self, hi, lo, depth=None) -> Pair:
        if getattr(self.ctx, "fused", False):
            return self.enc.renorm(hi, lo, level=depth)
        return self.enc.encode(self.enc.decode(hi, lo))

    def _xor_ct(self, a, b):
        return self.xor4.apply(a, b)

    def __call__(self, ct_hi, ct_lo, do_final_bootstrap: bool = True,
                 debug: Optional[Dict[str, Any]] = None) -> Pair:
        log = debug.__setitem__ if isinstance(debug, dict) else (lambda k, v: None)
        if getattr(self.ctx, "fused", False) and SHARE_GF_BASES and not isinstance(debug, dict):
            # 2x and 3x from one basis pair of x; 3*rot1(x) = rot1(3x); rot2, rot3 of x share one ModUp
            two, thr0 = self._gf_shared([2, 3], ct_hi, ct_lo)
            thr = self._rot_pair(thr0, 1)
            steps = [-4 * k * self.stride for k in (2, 3)]
            # rot2 / rot3 only feed XOR4s that run at XOR4_DEPTH: rotate the state there, not at its own level
            lo_hi, lo_lo = self.ctx.level_down(ct_hi, XOR4_DEPTH), self.ctx.level_down(ct_lo, XOR4_DEPTH)
            r2, r3 = zip(*self.ctx.pair_apply(self.ctx.rotate_many, (lo_hi, steps), (lo_lo, steps)))
        else:
            r1, r2, r3 = self._shifts(ct_hi, ct_lo)
            log("rotc1", r1), log("rotc2", r2), log("rotc3", r3), log("in", (ct_hi, ct_lo))
            two = self.gf_mult_2(ct_hi, ct_lo)
            thr = self.gf_mult_3(*r1)
        if getattr(self.ctx, "fused", False) and XOR_TREE and not isinstance(debug, dict):
            # XOR is associative: (2x ^ 3 rot1) and (rot2 ^ rot3) are independent, so they (and their renorms) run on two
            # lanes and only ONE XOR4 + renorm is left on the critical path after them (same three XOR4 pairs in total)
            def left():
                return self._renorm_pair(*self._xor_pair(two, thr), depth=XOR4_DEPTH)

            def right():
                return self._renorm_pair(*self._xor_pair(r2, r3), depth=XOR4_DEPTH)

            u, w = self.ctx.lane_map(lambda f: f(), [(left,), (right,)])
            acc = self._renorm_pair(*self._xor_pair(u, w), depth=0 if do_final_bootstrap else None)
            out_hi, out_lo = acc
            if do_final_bootstrap:
                out_hi, out_lo = self.ctx.bootstrap_pair(out_hi, out_lo, pre=self.ctx.to_intt)
            return out_hi, out_lo
        log("two", two), log("thr", thr)
        acc = self._xor_pair(two, thr)
        log("acc1", acc)
        acc = self._renorm_pair(*acc, depth=XOR4_DEPTH)
        acc = self._xor_pair(acc, r2)
        log("acc2", acc)
        acc = self._renorm_pair(*acc, depth=XOR4_DEPTH)
        acc = self._xor_pair(acc, r3)
        acc = self._renorm_pair(*acc, depth=0 if do_final_bootstrap else None)     # the bootstrap starts from level 0
        log("acc3", acc)
        out_hi, out_lo = acc
        if do_final_bootstrap:
            out_hi, out_lo = self.ctx.bootstrap_pair(out_hi, out_lo, pre=self.ctx.to_intt)
            log("out", (out_hi, out_lo))
        return out_hi, out_lo


class InvMixColumnsFHE(_MixBase):
    """out = 14*x xor 11*rot1(x) xor 13*rot2(x) xor 9*rot3(x)."""

    def __init__(self, ctx: EngineContext, xor4: XOR4LUT, use_hard_renorm: bool = True):
        super().__init__(ctx, xor4)
        self.stride = self.sc // 16
        self.enc = StateEncoder(ctx)
        self.use_hard_renorm = use_hard_renorm
        self._pt_row: List[Any] = []
        for r in range(4):
            m = np.zeros(self.sc, dtype=np.complex128)
            m[(r + 4 * np.arange(4)) * self.stride] = 1.0
            self._pt_row.append(ctx.encode(m))

    def gf_mult_9(self, h, l) -> Pair:
        return self._gf(9, h, l)

    def gf_mult_11(self, h, l) -> Pair:
        return self._gf(11, h, l)

    def gf_mult_13(self, h, l) -> Pair:
        return self._gf(13, h, l)

    def gf_mult_14(self, h, l) -> Pair:
        return self._gf(14, h, l)

    def _rot_rows_in_col(self, ct, k_rows: int):
        eng = self.ctx
        parts = [eng.rotate(eng.multiply(ct, m), k_rows * self.stride) for m in self._pt_row]
        out = eng.multiply(ct, 0.0)
        for p in parts:
            out = eng.add(out, p)
        return out

    def _renorm_pair(self, hi, lo, depth=None) -> Pair:
        if not self.use_hard_renorm:
            return hi, lo
        if getattr(self.ctx, "fused", False):
            return self.enc.renorm(hi, lo, level=depth)
        return self.enc.encode(self.enc.decode(hi, lo))

    def _xor(self, a, b):
        return self.xor4.apply(a, b)

    def __call__(self, ct_hi, ct_lo, do_final_bootstrap: bool = True,
                 debug: Optional[Dict[str, Any]] = None) -> Pair:
        log = debug.__setitem__ if debug is not None else (lambda k, v: None)
        if getattr(self.ctx, "fused", False) and SHARE_GF_BASES and debug is None:
            # all four multipliers from one basis pair of x; M_k(rot_k(x)) = rot_k(M_k(x))
            e14, m11, m13, m9 = self._gf_shared([14, 11, 13, 9], ct_hi, ct_lo)
            e11, e13, e9 = self.ctx.lane_map(self._rot_pair, [(m11, 1), (m13, 2), (m9, 3)])
        else:
            r1, r2, r3 = self._shifts(ct_hi, ct_lo)
            log("rotc1", r1), log("rotc2", r2), log("rotc3", r3)
            e14 = self.gf_mult_14(ct_hi, ct_lo); log("mul14", e14)
            e11 = self.gf_mult_11(*r1); log("mul11", e11)
            e13 = self.gf_mult_13(*r2); log("mul13", e13)
            e9 = self.gf_mult_9(*r3); log("mul9", e9)
        if getattr(self.ctx, "fused", False) and XOR_TREE and debug is None:
            def left():
                return self._renorm_pair(*self._xor_pair(e14, e11), depth=XOR4_DEPTH)

            def right():
                return self._renorm_pair(*self._xor_pair(e13, e9), depth=XOR4_DEPTH)

            u, w = self.ctx.lane_map(lambda f: f(), [(left,), (right,)])
            out_h, out_l = self._renorm_pair(*self._xor_pair(u, w), depth=0 if do_final_bootstrap else None)
            if do_final_bootstrap:
                out_h, out_l = self.ctx.bootstrap_pair(out_h, out_l)
            return out_h, out_l
        acc = self._xor_pair(e14, e11)
        log("acc1", acc)
        acc = self._renorm_pair(*acc, depth=XOR4_DEPTH)
        acc = self._xor_pair(acc, e13)
        log("acc2", acc)
        acc = self._renorm_pair(*acc, depth=XOR4_DEPTH)
        acc = self._xor_pair(acc, e9)
        out_h, out_l = self._renorm_pair(*acc, depth=0 if do_final_bootstrap else None)
        if do_final_bootstrap:
            out_h, out_l = self.ctx.bootstrap_pair(out_h, out_l)
        log("out", (out_h, out_l))
        return out_h, out_l
