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


class Zeta16NoiseReducer:
    def __init__(self, ctx: EngineContext, bootstrap_before: bool = False, bootstrap_after: bool = False):
        self.ctx = ctx
        self.alpha = 17.0 / 16.0
        self.beta = -1.0 / 16.0
        self.bootstrap_before = bootstrap_before
        self.bootstrap_after = bootstrap_after

    def _ensure_power_basis(self, ct: Any):
        eng = self.ctx
        try:
            return ct, eng.make_power_basis(ct, 8)
        except RuntimeError:                      # level / form trouble: one bootstrap, then retry (:24-29)
            ct = eng.bootstrap(ct)
            return ct, eng.make_power_basis(ct, 8)

    def apply(self, ct: Any) -> Any:
        eng = self.ctx
        x = eng.bootstrap(ct) if self.bootstrap_before else ct
        x, pos = self._ensure_power_basis(x)
        x1, x8 = pos[0], pos[7]
        x16 = eng.multiply(x8, x8)
        x17 = eng.multiply(x16, x1)
        y = eng.add(eng.multiply_plain(x1, self.alpha), eng.multiply_plain(x17, self.beta))
        return eng.bootstrap(y) if self.bootstrap_after else y

    def apply_pair(self, ct_hi: Any, ct_lo: Any) -> Tuple[Any, Any]:
        return self.apply(ct_hi), self.apply(ct_lo)


class Zeta16SnapNoMul:
    def __init__(self, ctx: EngineContext, bootstrap_before: bool = False, bootstrap_after: bool = False):
        self.ctx = ctx
        self.a = 9.0 / 8.0
        self.b = 1.0 / 8.0
        self.bootstrap_before = bootstrap_before
        self.bootstrap_after = bootstrap_after

    def _pb1_8(self, ct: Any):
        eng = self.ctx
        try:
            return eng.make_power_basis(ct, 8)
        except RuntimeError:
            return eng.make_power_basis(eng.bootstrap(ct), 8)

    def apply(self, ct: Any) -> Any:
        eng = self.ctx
        x = eng.bootstrap(ct) if self.bootstrap_before else ct
        pos = self._pb1_8(x)
        x1, x9 = pos[0], eng.conjugate(pos[6])
        y = eng.add(eng.multiply_plain(x1, self.a), eng.multiply_plain(x9, self.b))
        return eng.bootstrap(y) if self.bootstrap_after else y

    def apply_pair(self, hi: Any, lo: Any) -> Tuple[Any, Any]:
        return self.apply(hi), self.apply(lo)


class Zeta16Snap:
    def __init__(self, ctx: EngineContext, *, always_bs: bool = False):
        self.ctx = ctx
        self.always_bs = always_bs

    def _to_coeff(self, ct: Any) -> Any:
        try:
            return self.ctx.to_intt(ct)
        except Exception:
            return ct

    def _pb_1_8(self, ct: Any):
        eng = self.ctx
        ct = self._to_coeff(ct)
        if self.always_bs:
            ct = eng.bootstrap(ct)
        try:
            pos = eng.make_power_basis(ct, 8)
        except RuntimeError:
            ct = eng.bootstrap(ct)
            pos = eng.make_power_basis(ct, 8)
        return ct, pos

    def _mul_safe(self, a: Any, b: Any) -> Any:
        eng = self.ctx
        try:
            return eng.multiply(a, b)
        except RuntimeError:
            return eng.multiply(eng.bootstrap(a), b)

    def _scale_safe(self, ct: Any, s: float) -> Any:
        eng = self.ctx
        try:
            return eng.multiply(ct, float(s))
        except RuntimeError:
            return eng.multiply(eng.bootstrap(ct), float(s))

    def apply(self, ct: Any) -> Any:
        eng = self.ctx
        ct, pos = self._pb_1_8(ct)
        x1, x8 = pos[0], pos[7]
        x9 = eng.conjugate(pos[6])
        x17 = self._mul_safe(x9, x8)
        t1 = self._scale_safe(x1, 17.0 / 16.0)
        t2 = self._scale_safe(x17, 1.0 / 16.0)
        try:
            return eng.sub(t1, t2)
        except AttributeError:
            return eng.add(t1, eng.multiply(t2, -1.0))

    def apply_pair(self, hi: Any, lo: Any) -> Tuple[Any, Any]:
        return self.apply(hi), self.apply(lo)


class NoiseReducer:
    def __init__(self, ctx: EngineContext, n: int = 16, profile: bool = False):
        assert n >= 2
        self.ctx = ctx
        self.n = n
        self.alpha = 1.0 + 1.0 / n
        self.beta = -1.0 / n
        self.profile = profile
        self._last_stats: Optional[dict] = None

    def _ensure_read(self, ct: Any, deg: int = 1) -> Any:
        try:
            self.ctx.make_power_basis(ct, deg)
            return ct
        except RuntimeError:
            return self.ctx.bootstrap(ct)

    def _x_pow_nplus1(self, x: Any) -> Any:
        eng = self.ctx
        if self.n == 16:                              # the only branch the reference implements (:43-52)
            try:
                pos = eng.make_power_basis(x, 16)
            except RuntimeError:
                x = eng.bootstrap(x)
                pos = eng.make_power_basis(x, 16)
            return eng.relinearize(eng.multiply(pos[15], x))
        return None

    def apply(self, ct: Any) -> Any:
        eng = self.ctx
        t0 = time.perf_counter() if self.profile else None
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
    obj = json.loads(Path(json_path).read_text(encoding="utf-8"))
    max_k = max(int(k) for k, _, _ in obj["entries"])
    coeff = np.zeros(max_k + 1, dtype=np.complex128)
    for k, re, im in obj["entries"]:
        coeff[int(k)] = complex(re, im)
    return coeff


class Zeta16Snap1D:
    def __init__(self, ctx: EngineContext, coeff_1d: np.ndarray, bootstrap_before: bool = False):
        self.ctx = ctx
        self.sc = ctx.engine.slot_count
        self.coeff = np.asarray(coeff_1d, dtype=np.complex128)
        self.K = len(self.coeff) - 1
        self.bootstrap_before = bootstrap_before
        self.pt: Dict[int, Any] = {k: ctx.encode(np.full(self.sc, c, dtype=np.complex128))
                                   for k, c in enumerate(self.coeff) if abs(c) > 1e-12}

    def _power_basis_16(self, ct: Any) -> Dict[int, Any]:
        eng = self.ctx
        try:
            pos = eng.make_power_basis(ct, 8)
        except RuntimeError:
            ct = eng.bootstrap(ct)
            pos = eng.make_power_basis(ct, 8)
        zero_like = eng.multiply(ct, 0.0)
        try:
            basis0 = eng.add_plain(zero_like, 1.0)
        except RuntimeError:
            ct = eng.bootstrap(ct)
            basis0 = eng.add_plain(eng.multiply(ct, 0.0), 1.0)
        basis = {0: basis0}
        basis.update({k: pos[k - 1] for k in range(1, 9)})
        for k in range(9, 16):
            basis[k] = eng.conjugate(pos[(16 - k) - 1])
        return basis

    def apply(self, ct: Any) -> Any:
        eng = self.ctx
        if self.bootstrap_before:
            ct = eng.bootstrap(ct)
        if getattr(eng, "fused", False):
            # the same polynomial through the engine's fused linear combination: conj(x^k) = x^(16-k), one rescale
            pos = eng.make_power_basis(ct, 8)
            ks = [k for k in self.pt if k % 16 != 0]
            direct = [k for k in ks if k % 16 <= 8]
            mirror = [k for k in ks if k % 16 > 8]
            acc = eng.lincomb([pos[k % 16 - 1] for k in direct], [self.coeff[k] for k in direct]) if direct else None
            if mirror:
                m = eng.conjugate(eng.lincomb([pos[16 - k % 16 - 1] for k in mirror],
                                              [np.conj(self.coeff[k]) for k in mirror]))
                acc = m if acc is None else eng.add(acc, m)
            c0 = sum(self.coeff[k] for k in self.pt if k % 16 == 0)
            return eng.add_plain(acc, c0)
        basis = self._power_basis_16(ct)
        try:
            res = eng.multiply(ct, 0.0)
        except RuntimeError:
            ct = eng.bootstrap(ct)
            basis = self._power_basis_16(ct)
            res = eng.multiply(ct, 0.0)
        for k, pt in self.pt.items():
            res = eng.add(res, eng.multiply(basis[k % 16], pt))      # k > 15 folds with x^16 = 1 (:78-81)
        return res


class Zeta16SnapPair:
    def __init__(self, snap1d: Zeta16Snap1D):
        self.snap = snap1d

    def apply_pair(self, ct_hi: Any, ct_lo: Any) -> Tuple[Any, Any]:
        return self.snap.apply(ct_hi), self.snap.apply(ct_lo)
