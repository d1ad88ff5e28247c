"""CKKS bootstrapping on the CPU oracle.  TEST INFRASTRUCTURE ONLY (see oracle/ckks_oracle.py header).

Boundary restated: `desilofhe.Engine.bootstrap(ct, relin, conj, bootstrap_key)` as called from reference
`engine_context.py:147-162` (callers: `mixcol_final.py:158-163`, `invmixcolumns_fhe.py:166-168`).  The
reference's bootstrapping lives in the closed `desilofhe` wheel, so this restates the published full-slot
construction (Cheon-Han-Kim-Kim-Song EUROCRYPT'18; FFT-factored CoeffToSlot/SlotToCoeff with baby-step
giant-step of Chen-Chillotti-Song EUROCRYPT'19; cosine + double-angle EvalMod of Han-Ki CT-RSA'20) under
DESIGN.md spec S11, the same spec the CUDA engine's csrc/bootstrap.cu implements:

  1. level_down to level 0, ModRaise to level L: t = Delta_0 m + q_0 I  (|I| <= K)
  2. CoeffToSlot: `cts_groups` diagonal-sparse matrices (merged inverse special-FFT layers, no bit reversal)
     with the factor Delta_L / (2 q_0 K n) folded in; conjugation splits real and imaginary halves
  3. EvalMod on both halves: Chebyshev interpolant (degree d) of cos(2 pi (K x - 1/4) / 2^r) evaluated by
     Chebyshev division (baby steps T_1..T_m, giants T_2m, T_4m, ..), then r double-angle steps
  4. SlotToCoeff: `stc_groups` matrices (merged forward special-FFT layers) with q_0 / (2 pi Delta_0) folded in

Parity with the engine is by tolerance on decrypted slots (fp64 plan constants come from two different
libm call sites); every integer primitive underneath is checked bit-exactly elsewhere.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import os
import numpy as np
from numpy.polynomial import chebyshev as Cheb

from .ckks_oracle import Ct, OracleCKKS

Diags = Dict[int, np.ndarray]


# ---------------------------------------------------------------------------------------- plan (pure numpy)
def rot_group(n: int, M: int) -> np.ndarray:
    out = np.empty(n, dtype=np.int64)
    p = 1
    for j in range(n):
        out[j] = p
        p = p * 5 % M
    return out


def fft_layer(n: int, length: int, inverse: bool) -> Diags:
    """Diagonals {d: diag_d} (out[p] = sum_d diag_d[p] v[p+d]) of one special-FFT butterfly layer.
    forward  (ref_special_fft):  (u, x) -> (u + w x, u - w x)
    inverse  (ref_special_ifft): (a, b) -> ((a + b)/2, (a - b) conj(w) / 2)   [the 1/n is spread as 1/2 per layer]"""
    M = 4 * n
    lenh, lenq = length // 2, 4 * length
    gap = M // lenq
    rot = rot_group(n, M)
    p = np.arange(n)
    j = p % length % lenh
    first = (p % length) < lenh
    ang = 2 * np.pi * ((rot[j] % lenq) * gap) / M
    w = np.cos(ang) + 1j * np.sin(ang)
    d0 = np.zeros(n, dtype=np.complex128)
    dp = np.zeros(n, dtype=np.complex128)
    dm = np.zeros(n, dtype=np.complex128)
    if not inverse:
        d0[first], d0[~first] = 1, -w[~first]
        dp[first] = w[first]
        dm[~first] = 1
    else:
        d0[first], d0[~first] = 0.5, -0.5 * np.conj(w[~first])
        dp[first] = 0.5
        dm[~first] = 0.5 * np.conj(w[~first])
    out: Diags = {0: d0}
    for d, v in ((lenh % n, dp), ((-lenh) % n, dm)):
        out[d] = out[d] + v if d in out else v
    return out


def mat_mul(A: Diags, B: Diags, n: int) -> Diags:
    """(A B) as diagonals: diag_{a+b}[p] += A_a[p] * B_b[p+a]."""
    out: Diags = {}
    for a, da in A.items():
        for b, db in B.items():
            d = (a + b) % n
            t = da * np.roll(db, -a)
            out[d] = out[d] + t if d in out else t
    return {d: v for d, v in out.items() if np.abs(v).max() > 0}


def split_layers(nlayers: int, groups: int) -> List[int]:
    base, extra = divmod(nlayers, groups)
    return [base + (1 if i < extra else 0) for i in range(groups)]


def dft_plan(n: int, groups: int, inverse: bool, scale: float) -> List[Diags]:
    """Matrices in application order; `scale` is spread evenly over them."""
    lens = [2 << i for i in range(int(np.log2(n)))]                # 2, 4, .., n
    order = lens[::-1] if inverse else lens                         # CtS applies len = n first, StC len = 2 first
    mats: List[Diags] = []
    k = 0
    per = abs(scale) ** (1.0 / groups)
    for cnt in split_layers(len(order), groups):
        M = None
        for length in order[k:k + cnt]:
            Lm = fft_layer(n, length, inverse)
            M = Lm if M is None else mat_mul(Lm, M, n)             # later layer multiplies on the left
        k += cnt
        mats.append({d: v * per for d, v in M.items()})
    if scale < 0:
        mats[0] = {d: -v for d, v in mats[0].items()}
    return mats


def bsgs_split(diags: Diags, n: int) -> Tuple[int, int, Dict[int, Dict[int, np.ndarray]]]:
    """Signed baby-step/giant-step split of the diagonal set: d = stride*(j*n1 + i), 0 <= i < n1, j signed."""
    ds = sorted(diags)
    signed = [d if d <= n // 2 else d - n for d in ds]
    nz = [abs(d) for d in signed if d]
    stride = int(np.gcd.reduce(nz)) if nz else 1
    ks = [d // stride for d in signed]
    span = max(ks) - min(ks) + 1
    n1 = 1
    while n1 * n1 < span:
        n1 *= 2
    table: Dict[int, Dict[int, np.ndarray]] = {}
    for d, k in zip(ds, ks):
        j, i = divmod(k, n1)                                       # python divmod: i in [0, n1), j signed
        table.setdefault(j, {})[i] = diags[d]
    return stride, n1, table


# ---------------------------------------------------------------------------------------- oracle-side evaluation
class BootstrapOracle:
    def __init__(self, orc: OracleCKKS, K: int = 25, degree: int = 47, double_angle: int = 2, cts_groups: int = 3,
                 stc_groups: int = 3):
        self.o = orc
        self.K, self.degree, self.r = K, degree, double_angle
        n, L = orc.n, orc.L
        q0, S = float(orc.q[0]), orc.scales
        self.cts = dft_plan(n, cts_groups, True, S[L] / (2.0 * q0 * K))
        self.stc_groups = stc_groups
        f = lambda x: np.cos(2 * np.pi * (K * x - 0.25) / 2 ** double_angle)
        self.cheb = Cheb.chebinterpolate(f, degree)
        self.m = 1
        while self.m * self.m < degree + 1:
            self.m *= 2
        self.depth = cts_groups + (int(np.log2(self.m)) + 1 + self._giants()) + double_angle + stc_groups
        self.out_level = L - self.depth
        self.stc = dft_plan(n, stc_groups, False, q0 / (2.0 * np.pi * S[0]))

    def _giants(self) -> int:
        g, k = self.m, 0
        while g * 2 <= self.degree:
            g *= 2
            k += 1
        return k + 1 if self.degree >= self.m else 0

    # ---- primitives the oracle class lacks
    def mul_i(self, a: Ct, sign: int) -> Ct:
        o = self.o
        idx = o._idx_q(a.level)
        cp = np.array([(o.J[i] if sign >= 0 else o.moduli[i] - o.J[i]) for i in idx], dtype=np.uint64)
        cm = np.array([(o.moduli[i] - o.J[i] if sign >= 0 else o.J[i]) for i in idx], dtype=np.uint64)
        out = np.empty_like(a.c)
        for k in range(a.c.shape[0]):
            o.lib.ref_mul_const_batch(out[k], np.ascontiguousarray(a.c[k]), cp, cm, len(idx), o.N, o._mods(idx))
        return Ct(out, a.level, a.scale)

    def mod_raise(self, a: Ct) -> Ct:
        o = self.o
        a = o.level_down(a, 0)
        top = o._idx_q(o.L)
        polys = []
        for k in range(2):
            coef = o.intt(a.c[k][:1], [0])[0]
            q0 = np.uint64(o.q[0])
            signed = np.where(coef > q0 // np.uint64(2), coef.astype(np.int64) - np.int64(o.q[0]), coef.astype(np.int64))
            polys.append(o.ntt(o.reduce_i64(signed, top), top))
        return Ct(np.stack(polys), o.L, o.scales[o.L])

    def linear(self, a: Ct, diags: Diags) -> Ct:
        """One diagonal-sparse matrix by BSGS; rotleft(v, k) == rotate(ct, -k)."""
        o = self.o
        stride, n1, table = bsgs_split(diags, o.n)
        need = sorted({i for row in table.values() for i in row})
        babies = {i: (a if i == 0 else o.rotate(a, -i * stride)) for i in need}
        acc = None
        for j, row in sorted(table.items()):
            g = j * n1 * stride
            inner = None
            for i, dg in sorted(row.items()):
                t = o.mul_plain_vec(babies[i], np.roll(dg, g))     # rotleft(diag, -g)
                inner = t if inner is None else o.add_ct(inner, t)
            if g % o.n:
                inner = o.rotate(inner, -g)
            acc = inner if acc is None else o.add_ct(acc, inner)
        return acc

    def cheb_eval(self, x: Ct) -> Ct:
        o, m = self.o, self.m
        T: Dict[int, Ct] = {1: x}

        def get(k: int) -> Ct:
            if k not in T:
                a, b = (k + 1) // 2, k // 2                        # T_{a+b} = 2 T_a T_b - T_{a-b}
                prod = o.mul_ct(get(a), get(b))
                two = o.add_ct(prod, prod)
                T[k] = o.add_const(two, -1.0) if a == b else o.sub_ct(two, get(a - b))
            return T[k]

        def leaf(c: np.ndarray):
            acc = None
            for k in range(1, len(c)):
                if abs(c[k]) > 1e-300:
                    t = o.mul_const(get(k), complex(c[k]))
                    acc = t if acc is None else o.add_ct(acc, t)
            return acc, float(c[0])

        def rec(c: np.ndarray):
            """-> (ciphertext or None, pending constant)"""
            d = len(c) - 1
            if d < m:
                return leaf(c)
            g = m
            while g * 2 <= d:
                g *= 2
            # Chebyshev division by T_g: c = q T_g + r
            q = np.zeros(d - g + 1)
            r = np.array(c[:g], dtype=np.float64)
            q[0] = c[g]
            for k in range(g + 1, d + 1):
                q[k - g] = 2 * c[k]
                r[2 * g - k] -= c[k]
            qc, q0 = rec(q)
            rc, r0 = rec(r)
            Tg = get(g)
            # (q + q_0) T_g: the quotient's constant joins the quotient before the product (CKKS_CHEB_C0_FOLD=0: q_0 T_g as a
            # constant product of its own), in lockstep with csrc/bootstrap.cu cheb_eval
            folded = qc is not None and q0 != 0.0 and os.environ.get("CKKS_CHEB_C0_FOLD", "1") != "0"
            if folded:
                qc = o.add_const(qc, q0)
            t = o.mul_ct(qc, Tg) if qc is not None else None
            if q0 != 0.0 and not folded:
                t2 = o.mul_const(Tg, q0)
                t = t2 if t is None else o.add_ct(t, t2)
            if rc is not None:
                t = rc if t is None else o.add_ct(t, rc)
            return t, r0

        ct, c0 = rec(np.asarray(self.cheb, dtype=np.float64))
        return o.add_const(ct, c0)

    def eval_mod(self, x: Ct) -> Ct:
        o = self.o
        y = self.cheb_eval(x)
        for _ in range(self.r):
            sq = o.mul_ct(y, y)
            y = o.add_const(o.add_ct(sq, sq), -1.0)
        return y

    def bootstrap(self, a: Ct, debug: dict = None) -> Ct:
        o = self.o
        t = self.mod_raise(a)
        for M in self.cts:
            t = self.linear(t, M)
        cj = o.conjugate(t)
        re = o.add_ct(t, cj)
        im = self.mul_i(o.sub_ct(t, cj), -1)
        if debug is not None:
            debug["re"], debug["im"] = o.decrypt(re), o.decrypt(im)
        re, im = self.eval_mod(re), self.eval_mod(im)
        if debug is not None:
            debug["sre"], debug["sim"] = o.decrypt(re), o.decrypt(im)
        t = o.add_ct(re, self.mul_i(im, +1))
        for M in self.stc:
            t = self.linear(t, M)
        return t
