"""Homomorphic snap / noise reduction on zeta16 codewords: the polynomial alternatives to the hard
renorm, which needs the secret key on the evaluating side (reference `pipeline.py:65-69`).

Host-side mirrors of the reference classes, same names, constructor keywords, call sequence and recovery ladders:

  Zeta16NoiseReducer   zeta16_noise_reducter.py:6-57     f(x) = (17/16) x - (1/16) x^17, x^17 = (x^8)^2 x
  Zeta16SnapNoMul      zeta16_noise_reducter.py:60-106   f(x) = (9/8) x + (1/8) x^9,   x^9 = conj(x^7), no extra ct*ct
  Zeta16Snap           zeta16_noise_reducter.py:108-169  f(x) = (17/16) x - (1/16) x^17, x^17 = conj(x^7) x^8
  NoiseReducer         noise_reduction.py:14-79          the same map through make_power_basis(x, 16)
  Zeta16Snap1D / Pair  snapper_1d_z16.py:17-91           a 1-D LUT polynomial sum_k c_k x^k over the zeta16 basis

Every zeta16 codeword t is a fixed point of f with f'(t) = 0 (t^16 = 1), so a slot t (1 + e) comes back as
t (1 - O(e^2)): the error is squared, not removed -- unlike the hard renorm these maps cannot pull a slot back from
beyond the basin of its codeword.  `tests/test_snap.py` pins the call trace of each class to the unchanged reference
file on the slot stand-in and checks the contraction on the engine.
"""
from __future__ import annotations

import json
import time
from pathlib import Path
from typing import Any, Dict, Optional, Tuple

import numpy as np

from .context import EngineContext


def _with_bootstrap_retry(ctx, ct, fn):
    """The recovery ladder every reference class repeats: run `fn(ct)`; on a level / form `RuntimeError` bootstrap the
    ciphertext once and run it again.  Returns (ciphertext actually used, result)."""
    try:
        return ct, fn(ct)
    except RuntimeError:
        ct = ctx.bootstrap(ct)
        return ct, fn(ct)


class _PolySnap:
    """Shared shape of the three zeta16 maps: optional bootstrap before / after, positive powers x^1..x^8 first."""

    def __init__(self, ctx: EngineContext, bootstrap_before: bool = False, bootstrap_after: bool = False):
        self.ctx = ctx
        self.bootstrap_before = bootstrap_before
        self.bootstrap_after = bootstrap_after

    def _powers8(self, ct: Any):
        return _with_bootstrap_retry(self.ctx, ct, lambda c: self.ctx.make_power_basis(c, 8))

    def _combine(self, x: Any, pos) -> Any:
        raise NotImplementedError

    def apply(self, ct: Any) -> Any:
        x = self.ctx.bootstrap(ct) if self.bootstrap_before else ct
        x, pos = self._powers8(x)
        y = self._combine(x, pos)
        return self.ctx.bootstrap(y) if self.bootstrap_after else y

    def apply_pair(self, ct_hi: Any, ct_lo: Any) -> Tuple[Any, Any]:
        return self.apply(ct_hi), self.apply(ct_lo)


class Zeta16NoiseReducer(_PolySnap):
    """f(x) = (17/16) x - (1/16) x^17 with x^16 = (x^8)^2, x^17 = x^16 x   (zeta16_noise_reducter.py:31-52)."""

    def __init__(self, ctx: EngineContext, bootstrap_before: bool = False, bootstrap_after: bool = False):
        super().__init__(ctx, bootstrap_before, bootstrap_after)
        self.alpha, self.beta = 17.0 / 16.0, -1.0 / 16.0

    _ensure_power_basis = _PolySnap._powers8

    def _combine(self, x, pos):
        eng = self.ctx
        x17 = eng.multiply(eng.multiply(pos[7], pos[7]), pos[0])
        return eng.add(eng.multiply_plain(pos[0], self.alpha), eng.multiply_plain(x17, self.beta))


class Zeta16SnapNoMul(_PolySnap):
    """f(x) = (9/8) x + (1/8) x^9 with x^9 = conj(x^7): no product beyond the basis   (zeta16_noise_reducter.py:84-101)."""

    def __init__(self, ctx: EngineContext, bootstrap_before: bool = False, bootstrap_after: bool = False):
        super().__init__(ctx, bootstrap_before, bootstrap_after)
        self.a, self.b = 9.0 / 8.0, 1.0 / 8.0

    def _powers8(self, ct: Any):
        # the reference keeps the un-bootstrapped handle here (:76-82); only the basis matters afterwards
        return ct, _with_bootstrap_retry(self.ctx, ct, lambda c: self.ctx.make_power_basis(c, 8))[1]

    def _combine(self, x, pos):
        eng = self.ctx
        x9 = eng.conjugate(pos[6])
        return eng.add(eng.multiply_plain(pos[0], self.a), eng.multiply_plain(x9, self.b))


class Zeta16Snap(_PolySnap):
    """f(x) = (17/16) x - (1/16) x^17 with x^17 taken as conj(x^7) x^8 (one product; equal to x^17 on the unit circle
    only), every product guarded by its own bootstrap retry   (zeta16_noise_reducter.py:108-166)."""

    def __init__(self, ctx: EngineContext, *, always_bs: bool = False):
        super().__init__(ctx)
        self.always_bs = always_bs

    def _powers8(self, ct: Any):
        eng = self.ctx
        try:                                   # bootstrap wants coefficient form (:118-122)
            ct = eng.to_intt(ct)
        except Exception:
            pass
        if self.always_bs:
            ct = eng.bootstrap(ct)
        return _with_bootstrap_retry(eng, ct, lambda c: eng.make_power_basis(c, 8))

    def _guarded(self, a: Any, b) -> Any:
        return _with_bootstrap_retry(self.ctx, a, lambda c: self.ctx.multiply(c, b))[1]

    def _combine(self, x, pos):
        eng = self.ctx
        x17 = self._guarded(eng.conjugate(pos[6]), pos[7])
        t1 = self._guarded(pos[0], 17.0 / 16.0)
        t2 = self._guarded(x17, 1.0 / 16.0)
        if hasattr(eng, "sub"):
            return eng.sub(t1, t2)
        return eng.add(t1, eng.multiply(t2, -1.0))


class NoiseReducer:
    """f(x) = (1 + 1/n) x - (1/n) x^(n+1), n = 16, through make_power_basis(x, 16)   (noise_reduction.py:14-79)."""

    def __init__(self, ctx: EngineContext, n: int = 16, profile: bool = False):
        if n < 2:
            raise AssertionError("n must be at least 2")
        self.ctx, self.n, self.profile = ctx, n, profile
        self.alpha, self.beta = 1.0 + 1.0 / n, -1.0 / n
        self._last_stats: Optional[dict] = None

    def _ensure_read(self, ct: Any, deg: int = 1) -> Any:
        return _with_bootstrap_retry(self.ctx, ct, lambda c: self.ctx.make_power_basis(c, deg))[0]

    def _x_pow_nplus1(self, x: Any) -> Any:
        if self.n != 16:                        # the reference implements the nibble case only (:41-52)
            return None
        eng = self.ctx
        x, pos = _with_bootstrap_retry(eng, x, lambda c: eng.make_power_basis(c, 16))
        return eng.relinearize(eng.multiply(pos[15], x))

    def apply(self, ct: Any) -> Any:
        eng = self.ctx
        t0 = time.perf_counter()
        x = self._ensure_read(ct, 1)
        xn1 = self._x_pow_nplus1(x)
        y = eng.add(eng.multiply_plain(x, self.alpha), eng.multiply_plain(xn1, self.beta))
        if self.profile:
            self._last_stats = {"wall_s": time.perf_counter() - t0}
        return y

    def apply_pair(self, ct_hi: Any, ct_lo: Any) -> Tuple[Any, Any]:
        return self.apply(ct_hi), self.apply(ct_lo)

    def last_profile(self) -> Optional[dict]:
        return self._last_stats


def load_coeff1d(json_path: Path) -> np.ndarray:
    """Dense coefficient vector from the generator's sparse JSON ({"entries": [[k, re, im], ...]})."""
    entries = json.loads(Path(json_path).read_text(encoding="utf-8"))["entries"]
    coeff = np.zeros(1 + max(int(e[0]) for e in entries), dtype=np.complex128)
    for k, re, im in entries:
        coeff[int(k)] = re + 1j * im
    return coeff


class Zeta16Snap1D:
    """A 1-D LUT on one nibble ciphertext: sum_k c_k x^k over the zeta16 basis {1, x..x^8, conj(x^7)..conj(x)}
    (snapper_1d_z16.py:17-84); exponents above 15 fold with x^16 = 1."""

    def __init__(self, ctx: EngineContext, coeff_1d: np.ndarray, bootstrap_before: bool = False):
        self.ctx, self.sc, self.bootstrap_before = ctx, ctx.engine.slot_count, bootstrap_before
        self.coeff = np.asarray(coeff_1d, dtype=np.complex128)
        self.K = len(self.coeff) - 1
        nz = [k for k, c in enumerate(self.coeff) if abs(c) > 1e-12]
        self.pt: Dict[int, Any] = {k: ctx.encode(np.full(self.sc, self.coeff[k], dtype=np.complex128)) for k in nz}

    def _power_basis_16(self, ct: Any) -> Dict[int, Any]:
        eng = self.ctx
        ct, pos = _with_bootstrap_retry(eng, ct, lambda c: eng.make_power_basis(c, 8))
        zero = eng.multiply(ct, 0.0)
        try:
            one = eng.add_plain(zero, 1.0)
        except RuntimeError:
            one = eng.add_plain(eng.multiply(eng.bootstrap(ct), 0.0), 1.0)
        basis = dict(enumerate([one] + list(pos)))
        for k in range(9, 16):
            basis[k] = eng.conjugate(pos[15 - k])
        return basis

    def _apply_fused(self, ct: Any) -> Any:
        """The same polynomial through the engine's fused linear combination: one rescale, one conjugation."""
        eng = self.ctx
        pos = eng.make_power_basis(ct, 8)
        lo = [k for k in self.pt if 1 <= k % 16 <= 8]
        hi = [k for k in self.pt if k % 16 > 8]
        acc = eng.lincomb([pos[k % 16 - 1] for k in lo], [self.coeff[k] for k in lo]) if lo else None
        if hi:
            m = eng.conjugate(eng.lincomb([pos[15 - k % 16] for k in hi], [np.conj(self.coeff[k]) for k in hi]))
            acc = m if acc is None else eng.add(acc, m)
        return eng.add_plain(acc, sum(self.coeff[k] for k in self.pt if k % 16 == 0))

    def apply(self, ct: Any) -> Any:
        eng = self.ctx
        if self.bootstrap_before:
            ct = eng.bootstrap(ct)
        if getattr(eng, "fused", False):
            return self._apply_fused(ct)
        basis = self._power_basis_16(ct)
        try:
            res = eng.multiply(ct, 0.0)
        except RuntimeError:
            ct = eng.bootstrap(ct)
            basis, res = self._power_basis_16(ct), eng.multiply(ct, 0.0)
        for k, pt in self.pt.items():
            res = eng.add(res, eng.multiply(basis[k % 16], pt))
        return res


class Zeta16SnapPair:
    """The same 1-D snap on the hi and the lo nibble ciphertext (snapper_1d_z16.py:86-91)."""

    def __init__(self, snap1d: Zeta16Snap1D):
        self.snap = snap1d

    def apply_pair(self, ct_hi: Any, ct_lo: Any) -> Tuple[Any, Any]:
        return self.snap.apply(ct_hi), self.snap.apply(ct_lo)
