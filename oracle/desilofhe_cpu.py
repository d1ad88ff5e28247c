"""`desilofhe`-shaped surface over the CPU oracle.  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's backend (`from desilofhe import Engine, Ciphertext`, reference engine_context.py:1) is a closed wheel
that cannot run here.  This module gives the oracle port (oracle/ckks_oracle.py + ckks_ref.c, the same
RNS-CKKS arithmetic the CUDA engine is checked against bit for bit) the backend's class and method names, with the
calling conventions of SURVEY.md Appendix A, so that the reference's call sequence -- the unchanged reference modules
through tests/refload.py, or the host mirror with `fused=False` -- can be EXECUTED on the host CPU:

  * tests compare engine results with it call for call;
  * `bench.py --impl reference` and the `cpu_baseline` leg time it ("kind": "port"; SURVEY.md 8d (ii): "same op sequence
    through the same shim, mode='cpu'").

Product code never imports this module (the product path fails loudly without the CUDA library).
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np

from .ckks_oracle import Ct as _OCt
from .ckks_oracle import OracleCKKS
from .params import make_params

__all__ = ["Engine", "Ciphertext", "Plaintext"]

COUNTS = {"mul_cc": 0, "conj": 0, "rot": 0, "mul_pt": 0, "add": 0, "boot": 0, "enc": 0, "dec": 0}


class Ciphertext:
    __slots__ = ("ct", "_low", "ntt_form")

    def __init__(self, ct: _OCt):
        self.ct = ct
        self._low = {}          # memoised level alignments (the CUDA engine memoises them too)
        self.ntt_form = True

    @property
    def level(self) -> int:
        return self.ct.level

    @property
    def batch(self) -> int:
        return 1


class Plaintext:
    def __init__(self, vec: np.ndarray, n: int):
        v = np.asarray(vec)
        first = v.flat[0] if v.size else 0.0
        self.const = complex(first) if (v.size == n and np.all(v == first)) else None
        self.vec = None if self.const is not None else np.concatenate([v, np.zeros(n - v.size)]).astype(np.complex128)


class _Key:
    pass


class Engine:
    """Keyword surface of the B200 drop-in (`desilofhe.Engine` in this repository) on the CPU oracle."""

    def __init__(self, *, mode: str = "cpu", use_bootstrap: bool = False, use_multiparty: bool = False,
                 thread_count: Optional[int] = None, device_id: int = 0, max_level: Optional[int] = None,
                 seed: int = 1, logn: int = 16, levels: int = 21, dnum: int = 3, hamming_weight: int = 192,
                 fresh_level: int = -1, q0_bits: int = 50, **_ignored):
        if thread_count:
            os.environ["OMP_NUM_THREADS"] = str(int(thread_count))
        if max_level is not None:
            levels = int(max_level)
        if fresh_level < 0:
            fresh_level = 14 if (use_bootstrap and levels > 14) else levels
        self.prm = make_params(logn=logn, levels=levels, dnum=dnum, hamming_weight=hamming_weight, fresh_level=fresh_level,
                               q0_bits=q0_bits)
        self.orc = OracleCKKS(self.prm, seed=seed)
        self.slot_count = self.orc.n
        self.use_bootstrap = use_bootstrap
        self._boot = None
        self.backend = "oracle port (CPU)"

    # ---- keys (engine_context.py:44-50)
    def create_secret_key(self):
        self.orc.keygen_secret()
        return _Key()

    def create_public_key(self, sk):
        self.orc.keygen_public()
        return _Key()

    def create_relinearization_key(self, sk):
        self.orc.keygen_relin()
        return _Key()

    def create_conjugation_key(self, sk):
        self.orc.keygen_galois(self.orc.galois_conj())
        return _Key()

    def create_rotation_key(self, sk, steps=None):
        return _Key()                      # per-step Galois keys are derived on first use, as in the engine

    def create_bootstrap_key(self, sk):
        if self.use_bootstrap and self._boot is None:
            from .bootstrap_oracle import BootstrapOracle
            self._boot = BootstrapOracle(self.orc, K=25, degree=47, double_angle=3)
        return _Key()

    # ---- data movement (engine_context.py:56-63)
    def encode(self, vec) -> Plaintext:
        return Plaintext(vec, self.slot_count)

    def _slots(self, data) -> np.ndarray:
        v = np.asarray(data, dtype=np.complex128)
        out = np.zeros(self.slot_count, dtype=np.complex128)
        out[:v.size] = v
        return out

    def encrypt(self, data, pk=None, level: int = -1) -> Ciphertext:
        COUNTS["enc"] += 1
        return Ciphertext(self.orc.encrypt(self._slots(data), None if level < 0 else level))

    def decrypt(self, ct: Ciphertext, sk=None) -> np.ndarray:
        COUNTS["dec"] += 1
        return self.orc.decrypt(ct.ct)

    # ---- arithmetic (engine_context.py:65-98); binary operations align levels themselves
    def _down(self, c: Ciphertext, level: int) -> _OCt:
        if c.ct.level == level:
            return c.ct
        if level not in c._low:
            c._low[level] = self.orc.level_down(c.ct, level)
        return c._low[level]

    def _pair(self, a: Ciphertext, b: Ciphertext):
        l = min(a.level, b.level)
        return self._down(a, l), self._down(b, l)

    @staticmethod
    def _is_scalar(x) -> bool:
        return isinstance(x, (int, float, complex, np.integer, np.floating, np.complexfloating))

    def multiply(self, a, b, relin=None) -> Ciphertext:
        if isinstance(b, Ciphertext) and not isinstance(a, Ciphertext):
            a, b = b, a
        if isinstance(b, Ciphertext):
            x, y = self._pair(a, b)
            if relin is not None:
                COUNTS["mul_cc"] += 1
                return Ciphertext(self.orc.mul_ct(x, y))
            return Ciphertext(self.orc.rescale(self.orc.tensor(x, y)))
        COUNTS["mul_pt"] += 1
        if isinstance(b, Plaintext):
            if b.const is not None:
                return Ciphertext(self.orc.mul_const(a.ct, b.const))
            return Ciphertext(self.orc.mul_plain_vec(a.ct, b.vec))
        if self._is_scalar(b):
            return Ciphertext(self.orc.mul_const(a.ct, complex(b)))
        return self.multiply(a, self.encode(b))

    def add(self, a, b) -> Ciphertext:
        if isinstance(b, Ciphertext) and not isinstance(a, Ciphertext):
            a, b = b, a
        COUNTS["add"] += 1
        if isinstance(b, Ciphertext):
            x, y = self._pair(a, b)
            return Ciphertext(self.orc.add_ct(x, y))
        if isinstance(b, Plaintext):
            if b.const is not None:
                return Ciphertext(self.orc.add_const(a.ct, b.const))
            return Ciphertext(self.orc.add_plain_vec(a.ct, b.vec))
        if self._is_scalar(b):
            return Ciphertext(self.orc.add_const(a.ct, complex(b)))
        return self.add(a, self.encode(b))

    def subtract(self, a, b) -> Ciphertext:
        COUNTS["add"] += 1
        if isinstance(a, Ciphertext) and isinstance(b, Ciphertext):
            x, y = self._pair(a, b)
            return Ciphertext(self.orc.sub_ct(x, y))
        if isinstance(a, Ciphertext):
            if isinstance(b, Plaintext):
                b = b.const if b.const is not None else b.vec
            return self.add(a, -complex(b) if self._is_scalar(b) else -np.asarray(b))
        raise TypeError("subtract(plain, ciphertext) is not used by the AES path")

    def add_plain(self, ct: Ciphertext, val) -> Ciphertext:
        return self.add(ct, val)

    def make_power_basis(self, ct: Ciphertext, degree: int, relin=None) -> List[Ciphertext]:
        COUNTS["mul_cc"] += max(int(degree) - 1, 0)
        return [Ciphertext(c) for c in self.orc.power_basis(ct.ct, int(degree))]

    def conjugate(self, ct: Ciphertext, key=None) -> Ciphertext:
        COUNTS["conj"] += 1
        return Ciphertext(self.orc.conjugate(ct.ct))

    def rotate(self, ct: Ciphertext, key, steps: int) -> Ciphertext:
        COUNTS["rot"] += 1
        return Ciphertext(self.orc.rotate(ct.ct, int(steps)))

    def relinearize(self, ct: Ciphertext, relin=None) -> Ciphertext:
        if ct.ct.c.shape[0] != 3:
            raise RuntimeError("ciphertext should have 3 polynomials")
        return Ciphertext(self.orc.relinearize(ct.ct))

    def bootstrap(self, ct: Ciphertext, relin=None, conj=None, bsk=None) -> Ciphertext:
        if self._boot is None:
            raise RuntimeError("engine was created without use_bootstrap=True")
        COUNTS["boot"] += 1
        return Ciphertext(self._boot.bootstrap(ct.ct))

    def ntt(self, x):
        return x

    def intt(self, x):
        return x
