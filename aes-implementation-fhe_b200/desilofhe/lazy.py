"""Deferred evaluation behind the reference's call-for-call API.

The reference's step classes evaluate every look-up-table polynomial one term at a time through the backend's public
methods (reference `xor4_lut.py:63-74`, `mixcol_final.py:80-99`, `invmixcolumns_fhe.py:76-90`, `sub_bytes_lut.py:46-73`):

    term = engine.multiply(A[p], B[q], relin_key)      # ct x ct: key switch + rescale
    term = engine.multiply(term, constant_plaintext)   # rescale
    res  = engine.add(res, term)

and rebuild identical power bases (`make_power_basis(ct, 8)` + 7 conjugations) for every table.  Issued one for one this
costs 1 588 key switches per AES round; the same polynomials evaluated by the engine's fused kernels (`ckks_lut2`,
`ckks_lincomb`, csrc/lut.cu) cost a third of that.  To give the UNCHANGED callers the fused path, an `Engine(lazy=True)`
does not execute those calls at once: it returns ciphertext handles that stand for small expressions

    LPow(base, k)      element k of make_power_basis(base, d)      -- computed on demand, once per (base, k)
    LConj(x)           conjugate(x)                                -- once per x
    LProd(a, b)        multiply(a, b, relin_key)
    LSum               c0 + sum c_t * a_t * b_t + sum c_t * x_t + sum y_t   (what the add/multiply chains above build)

and evaluates a sum when something needs its value (decrypt, rotate, bootstrap, a product with another ciphertext, ...):
the bilinear terms as ONE fused LUT (one relinearisation), the linear terms as one fused linear combination, the terms
on conjugates as conj(sum conj(c) x) -- one conjugation instead of one per term.  The value differs from the call-for-call
result only by rounding noise (fewer rescales); decoded bytes are identical (tests/test_lazy_engine.py,
tests/test_reference_on_engine.py).  Level bookkeeping follows the reference's semantics exactly (every multiply consumes
one level), and `make_power_basis` still raises its "level should be positive" RuntimeError AT THE CALL,
which the reference's recovery ladders depend on (xor4_lut.py:33-51).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple


class LPow:
    __slots__ = ("base", "k")

    def __init__(self, base, k: int):
        self.base, self.k = base, k


class LConj:
    __slots__ = ("x",)

    def __init__(self, x):
        self.x = x


class LProd:
    __slots__ = ("a", "b")

    def __init__(self, a, b):
        self.a, self.b = a, b


class LSum:
    """c0 + sum_t c_t a_t b_t + sum_t c_t x_t + sum_t y_t.  `zero_src`: how to make a zero ciphertext of the right shape
    when the sum has no ciphertext term at all (multiply(ct, 0.0) / subtract(ct, ct) in the callers)."""
    __slots__ = ("t2", "t1", "plain", "c0", "zero_src")

    def __init__(self, t2=None, t1=None, plain=None, c0: complex = 0j, zero_src=None):
        self.t2: List[Tuple[complex, object, object]] = list(t2 or [])
        self.t1: List[Tuple[complex, object]] = list(t1 or [])
        self.plain: List[object] = list(plain or [])
        self.c0 = complex(c0)
        self.zero_src = zero_src          # (ciphertext, levels_below_it)

    def merged(self, other: "LSum") -> "LSum":
        return LSum(self.t2 + other.t2, self.t1 + other.t1, self.plain + other.plain, self.c0 + other.c0,
                    self.zero_src or other.zero_src)


def depth_of(degree: int) -> int:
    d = 0
    while (1 << d) < degree:
        d += 1
    return d


class LazyOps:
    """Mixin of `desilofhe.Engine`: builds and evaluates the deferred expressions.  `self._wrap(lazy, level, batch)` makes a
    deferred Ciphertext, `ct._lz` is its expression (None once it has a value), `ct._h` forces it."""

    # ------------------------------------------------------------------ construction
    def _is_atom(self, ct) -> bool:
        return ct._lz is None or isinstance(ct._lz, (LPow, LConj))

    def _lazy_power_basis(self, ct, degree: int):
        need = depth_of(degree)
        if ct.level < need:                 # same error, same moment as the eager call (xor4_lut.py:33-51 matches on it)
            raise RuntimeError(f"ciphertext level should be positive: make_power_basis: ciphertext level should be positive "
                               f"for multiplication (level {ct.level}, need {need})")
        return [self._wrap(LPow(ct, k), ct.level - depth_of(k), ct.batch) for k in range(1, degree + 1)]

    def _lazy_conjugate(self, ct):
        memo = ct._memo
        hit = memo.get("lconj")
        if hit is None:
            hit = memo["lconj"] = self._wrap(LConj(ct), ct.level, ct.batch)
        return hit

    def _lazy_multiply_cc(self, a, b):
        lvl = min(a.level, b.level)
        if lvl < 1:
            raise RuntimeError("ciphertext level should be positive: multiply: ciphertext level should be positive for "
                               f"multiplication (level {lvl}, need 1)")
        return self._wrap(LProd(a, b), lvl - 1, max(a.batch, b.batch))

    def _lazy_multiply_const(self, a, c: complex):
        if a.level < 1:
            raise RuntimeError("ciphertext level should be positive: multiply: ciphertext level should be positive for "
                               f"multiplication (level {a.level}, need 1)")
        if c == 0:
            return self._wrap(LSum(zero_src=(a, 1)), a.level - 1, a.batch)
        z = a._lz
        if isinstance(z, LProd):
            return self._wrap(LSum(t2=[(c, z.a, z.b)]), a.level - 1, a.batch)
        if self._is_atom(a):
            return self._wrap(LSum(t1=[(c, a)]), a.level - 1, a.batch)
        return None                         # a sum times a constant: evaluated, then multiplied eagerly

    def _as_sum(self, ct) -> Optional[LSum]:
        z = ct._lz
        if isinstance(z, LSum):
            return z
        if self._is_atom(ct):
            return LSum(plain=[ct])
        return None

    def _lazy_add(self, a, b):
        sa, sb = self._as_sum(a), self._as_sum(b)
        if sa is None or sb is None or (a._lz is None and b._lz is None):
            return None
        return self._wrap(sa.merged(sb), min(a.level, b.level), max(a.batch, b.batch))

    def _lazy_add_const(self, a, c: complex):
        z = a._lz
        if isinstance(z, LSum):
            return self._wrap(LSum(z.t2, z.t1, z.plain, z.c0 + c, z.zero_src), a.level, a.batch)
        return None

    def _lazy_sub_self(self, a):
        return self._wrap(LSum(zero_src=(a, 0)), a.level, a.batch)

    # ------------------------------------------------------------------ evaluation
    def _force(self, ct) -> None:
        """Give `ct` a value: afterwards ct._hraw is a live handle and ct._lz is None."""
        z = ct._lz
        if z is None:
            return
        if isinstance(z, LPow):
            val = self._pow_value(z.base, z.k)
        elif isinstance(z, LConj):
            val = self._conj_value(z.x)
        elif isinstance(z, LProd):
            val = self._eager_multiply_cc(z.a, z.b)
        else:
            val = self._sum_value(z, ct.level)
        # adopt the value's handle (the deferred handle object is what the caller holds)
        ct._hraw, val._hraw = val._hraw, 0
        ct._lz = None

    def _pow_value(self, base, k: int):
        """base^k with the same product tree as make_power_basis (k = k//2 + (k+1)//2): bit-identical to the full basis."""
        if k == 1:
            return self._eager_copy(base)
        memo: Dict = base._memo
        hit = memo.get(("pow", k))
        if hit is not None:
            return self._eager_copy(hit)
        self._materialize_powers(base, [k])
        return self._eager_copy(memo[("pow", k)])

    def _materialize_powers(self, base, ks) -> None:
        memo: Dict = base._memo
        todo = sorted({k for k in ks if k > 1 and ("pow", k) not in memo})
        if not todo:
            return
        got = self._eager_power_basis_sparse(base, max(todo), todo)       # one engine call: lanes per generation inside
        for k, c in enumerate(got, start=1):
            if c is not None and k > 1 and ("pow", k) not in memo:
                memo[("pow", k)] = c

    def _atom_value(self, ct):
        """A Ciphertext WITH a value for an atom or any other deferred handle (shared through the memo tables)."""
        z = ct._lz
        if z is None:
            return ct
        if isinstance(z, LPow):
            if z.k == 1:
                return self._atom_value(z.base)
            self._materialize_powers(z.base, [z.k])
            return z.base._memo[("pow", z.k)]
        if isinstance(z, LConj):
            inner = self._atom_value(z.x)
            hit = inner._memo.get("conj")
            if hit is None:
                hit = inner._memo["conj"] = self._eager_conjugate(inner)
            return hit
        self._force(ct)
        return ct

    def _conj_value(self, x):
        return self._eager_copy(self._atom_value(self._lazy_conjugate(x)))

    def _prepare_atoms(self, atoms) -> None:
        """Compute everything the listed atoms need, grouped so that independent work shares engine calls / stream lanes:
        all powers of one base in one sparse power-basis call, all conjugations side by side."""
        by_base: Dict[int, Tuple[object, set]] = {}
        conj: List[object] = []
        seen = set()
        for a in atoms:
            if id(a) in seen:
                continue
            seen.add(id(a))
            z = a._lz
            if isinstance(z, LConj):
                conj.append(a)
                z = z.x._lz
            if isinstance(z, LPow) and z.k > 1:
                by_base.setdefault(id(z.base), (z.base, set()))[1].add(z.k)
        bases = list(by_base.values())
        for b, _ in bases:
            if b._lz is not None:
                self._force(b)
        if len(bases) > 1:
            self.lane_map(lambda b, ks: self._materialize_powers(b, ks), [(b, sorted(ks)) for b, ks in bases])
        elif bases:
            self._materialize_powers(bases[0][0], sorted(bases[0][1]))
        inners = []
        for a in conj:
            inner = self._atom_value(a._lz.x)
            if "conj" not in inner._memo and all(inner is not i for i in inners):
                inners.append(inner)
        if len(inners) > 1:
            for inner, val in zip(inners, self.lane_map(self._eager_conjugate, [(i,) for i in inners])):
                inner._memo["conj"] = val
        elif inners:
            inners[0]._memo["conj"] = self._eager_conjugate(inners[0])

    BSGS_MIN_DEGREE = 32           # below this the plain power basis is as cheap

    def _single_base(self, terms):
        """The common base if every term is c * base^k (distinct k) and the degree is high enough for baby-step/giant-step."""
        base, ks = None, set()
        for _, x in terms:
            z = x._lz
            if not isinstance(z, LPow) or (base is not None and z.base is not base) or z.k in ks:
                return None
            base = z.base
            ks.add(z.k)
        if base is None or max(ks) <= self.BSGS_MIN_DEGREE or len(ks) < self.BSGS_MIN_DEGREE or max(ks) > 256:
            return None
        return base

    def _poly_bsgs(self, base, coef: Dict[int, complex]):
        """sum_k coef[k] base^k (k >= 1) as sum_j (sum_i coef[16 j + i] base^i) * (base^16)^j: the inner sums are fused
        linear combinations of the 15 baby powers (no key switch), the outer products are accumulated as 3-polynomial
        ciphertexts and relinearised ONCE.  The inner sums sit one level below the babies, i.e. above every giant power
        from (base^16)^2 on, so the result is at the level the call-for-call evaluation reports (depth(max k) + 1)."""
        bval = self._atom_value(base)
        kmax = max(coef)
        J = kmax // 16
        self._materialize_powers(bval, list(range(2, 17)))
        memo = bval._memo
        baby = {1: bval}
        baby.update({i: memo[("pow", i)] for i in range(2, 16)})
        G = memo[("pow", 16)]
        if J > 1:
            need = [j for j in range(2, J + 1) if any((16 * j + i) in coef for i in range(16))]
            self._materialize_powers(G, need)
        giant = {1: G}
        giant.update({j: G._memo[("pow", j)] for j in range(2, J + 1) if ("pow", j) in G._memo})
        acc3, acc2 = None, None
        for j in range(0, J + 1):
            idx = [i for i in range(1, 16) if (16 * j + i) in coef]
            c0 = coef.get(16 * j, 0j) if j else 0j
            if not idx and c0 == 0:
                continue
            if idx:
                u = self._eager_lincomb([baby[i] for i in idx], [coef[16 * j + i] for i in idx])
                if c0 != 0:
                    u = self._eager_add_const(u, c0)
            if j == 0:
                acc2 = u
                continue
            if idx:
                t = self._new(self._lib.ckks_mul_norelin, u._h, giant[j]._h)
                acc3 = t if acc3 is None else self._eager_add(acc3, t)
            else:                                        # only the pure giant power: c * (base^16)^j
                t = self._new(self._lib.ckks_mul_const, giant[j]._h, c0.real, c0.imag)
                acc2 = t if acc2 is None else self._eager_add(acc2, t)
        out = None
        if acc3 is not None:
            out = self._new(self._lib.ckks_relinearize, acc3._h)
        if acc2 is not None:
            out = acc2 if out is None else self._eager_add(out, acc2)
        return out

    def _sum_value(self, S: LSum, level: int):
        parts = []
        # ---- bilinear terms: fused LUT(s), at most 16 distinct left and 16 distinct right operands each
        if S.t2:
            self._prepare_atoms([x for _, a, b in S.t2 for x in (a, b)])
            groups: List[Tuple[List, List, List]] = []
            for c, a, b in S.t2:
                va, vb = self._atom_value(a), self._atom_value(b)
                for A, B, T in groups:
                    ia = next((i for i, x in enumerate(A) if x is va), None)
                    ib = next((i for i, x in enumerate(B) if x is vb), None)
                    if (ia is not None or len(A) < 16) and (ib is not None or len(B) < 16):
                        if ia is None:
                            A.append(va); ia = len(A) - 1
                        if ib is None:
                            B.append(vb); ib = len(B) - 1
                        T.append((ia, ib, c))
                        break
                else:
                    groups.append(([va], [vb], [(0, 0, c)]))
            for A, B, T in groups:
                n = max(len(A), len(B))
                parts.append(self._eager_lut2(A + [None] * (n - len(A)), B + [None] * (n - len(B)), T))
        # ---- linear terms: sum c x + conj(sum conj(c) x') for the terms on conjugates (ONE conjugation)
        if S.t1:
            direct = [(c, x) for c, x in S.t1 if not isinstance(x._lz, LConj)]
            mirror = [(c.conjugate(), x._lz.x) for c, x in S.t1 if isinstance(x._lz, LConj)]
            bd = self._single_base(direct) if direct else None
            bm = self._single_base(mirror) if mirror else None
            base = bd if bd is not None else bm
            if base is not None and (not direct or bd is base) and (not mirror or bm is base):
                # a polynomial of high degree in ONE base (the S-box polynomials of sub_bytes_lut.py:63-71): baby-step /
                # giant-step instead of the full power basis -- 22 products for degree 128 instead of 127
                if direct:
                    parts.append(self._poly_bsgs(base, {x._lz.k: c for c, x in direct}))
                if mirror:
                    parts.append(self._eager_conjugate(self._poly_bsgs(base, {x._lz.k: c for c, x in mirror})))
            else:
                self._prepare_atoms([x for _, x in direct] + [x for _, x in mirror])
                if direct:
                    parts.append(self._eager_lincomb([self._atom_value(x) for _, x in direct], [c for c, _ in direct]))
                if mirror:
                    parts.append(self._eager_conjugate(self._eager_lincomb([self._atom_value(x) for _, x in mirror],
                                                                           [c for c, _ in mirror])))
        shared = set()
        for y in S.plain:
            v = self._atom_value(y)
            shared.add(id(v))
            parts.append(v)
        if not parts:
            src, below = S.zero_src
            acc = self._eager_zero(self._atom_value(src), below)
        else:
            acc = parts[0]
            for p in parts[1:]:
                acc = self._eager_add(acc, p)
            if id(acc) in shared:
                acc = self._eager_copy(acc)          # never hand out an operand's own handle
        if S.c0 != 0:
            acc = self._eager_add_const(acc, S.c0)
        if acc.level > level:                          # a sum of high-level terms only: the callers' level bookkeeping wins
            acc = self._eager_level_down(acc, level)
        return acc
