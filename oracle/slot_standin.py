"""Tier-A oracle: slot-domain stand-in for the CKKS backend.  TEST INFRASTRUCTURE ONLY.

Nothing in the product path may import this file (only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg).  It restates the *observable slot semantics* the
reference callers rely on at the `desilofhe` boundary (reference
`engine_context.py:44-204`; behavioural inferences listed in SURVEY.md Appendix A and
rebuilt as described in SURVEY.md Appendix D): a ciphertext is a plain complex128
vector plus a level counter, every op is exact fp64 slot arithmetic, `multiply`
burns one level, `rotate(ct, +r) == np.roll(v, +r)`, `conjugate == np.conj`.

It also records an operation trace and per-op counters, used to (i) pin the op-count
table of SURVEY.md Appendix B and (ii) prove that our host-side mirror of the
reference AES modules issues the *same* engine-call sequence as the reference files.

parity: pinned at decoded-byte level against FIPS-197 tables embedded in the reference
(`sub_bytes_lut.py:86-103`) and the Appendix-C golden byte fixtures; there is no
ciphertext-level golden data in the reference.
"""
from __future__ import annotations

import hashlib
import os
from collections import Counter
from typing import List

import numpy as np

SLOT_COUNT = int(os.environ.get("STANDIN_SLOTS", "32768"))
FRESH_LEVEL = int(os.environ.get("STANDIN_FRESH_LEVEL", "29"))
BOOT_LEVEL = int(os.environ.get("STANDIN_BOOT_LEVEL", "14"))


class Ciphertext:
    __slots__ = ("v", "level", "uid", "ntt")

    def __init__(self, v, level, uid, ntt=True):
        self.v = v
        self.level = level
        self.uid = uid
        self.ntt = ntt


class Plaintext:
    __slots__ = ("v", "uid")

    def __init__(self, v, uid):
        self.v = v
        self.uid = uid


class _Key:
    def __init__(self, kind):
        self.kind = kind


class Engine:
    """Same constructor keywords and method names as the closed `desilofhe.Engine`
    as used from reference `engine_context.py:17-50`."""

    def __init__(self, *, max_level=None, mode="cpu", use_bootstrap=False, use_multiparty=False,
                 thread_count=1, device_id=0, slot_count=None, fresh_level=None, boot_level=None,
                 noise_std=0.0, seed=0):
        self.slot_count = int(slot_count or SLOT_COUNT)
        self.fresh_level = int(fresh_level if fresh_level is not None else
                               (max_level if max_level is not None else FRESH_LEVEL))
        self.boot_level = int(boot_level if boot_level is not None else BOOT_LEVEL)
        self.noise_std = float(noise_std)
        self._rng = np.random.default_rng(seed)
        self.counters: Counter = Counter()
        self.trace: List[tuple] = []
        self.trace_enabled = False
        self._uid = 0
        self.stage = ""

    # -- bookkeeping ---------------------------------------------------------
    def _new(self, v, level, ntt=True):
        self._uid += 1
        if self.noise_std:
            v = v + self.noise_std * (self._rng.standard_normal(v.shape) + 1j * self._rng.standard_normal(v.shape))
        return Ciphertext(v, level, self._uid, ntt)

    def _rec(self, op, *args):
        self.counters[(self.stage, op)] += 1
        if self.trace_enabled:
            self.trace.append((op,) + args)

    def trace_digest(self) -> str:
        h = hashlib.sha256()
        for t in self.trace:
            h.update(repr(t).encode())
        return h.hexdigest()

    def reset_trace(self):
        self.trace.clear()
        self.counters.clear()
        self._uid = 0

    @staticmethod
    def _vec_tag(v: np.ndarray):
        """Short content tag for a plaintext vector (constant vectors collapse to the constant)."""
        v = np.asarray(v)
        if v.size and np.all(v == v.flat[0]):
            c = complex(v.flat[0])
            return ("const", round(c.real, 12) + 0.0, round(c.imag, 12) + 0.0)   # +0.0 folds -0.0
        return ("vec", hashlib.sha256(np.ascontiguousarray(v.astype(np.complex128)).tobytes()).hexdigest()[:16])

    # -- keys ------------------------------------------------------------------
    def create_secret_key(self):
        return _Key("sk")

    def create_public_key(self, sk):
        return _Key("pk")

    def create_relinearization_key(self, sk):
        return _Key("relin")

    def create_conjugation_key(self, sk):
        return _Key("conj")

    def create_rotation_key(self, sk):
        return _Key("rot")

    def create_bootstrap_key(self, sk):
        return _Key("boot")

    # -- data movement ---------------------------------------------------------
    def encode(self, vec):
        v = np.asarray(vec)
        v = v.astype(np.complex128) if np.iscomplexobj(v) else v.astype(np.float64).astype(np.complex128)
        if v.shape != (self.slot_count,):
            raise ValueError("encode expects a full slot vector")
        self._uid += 1
        self._rec("encode", self._vec_tag(v))
        return Plaintext(v.copy(), self._uid)

    def encrypt(self, data, pk):
        v = np.asarray(data, dtype=np.complex128)
        if v.shape != (self.slot_count,):
            raise ValueError("encrypt expects a full slot vector")
        out = self._new(v.copy(), self.fresh_level)
        self._rec("encrypt", self._vec_tag(v), out.uid)
        return out

    def decrypt(self, ct, sk):
        self._rec("decrypt", ct.uid)
        return ct.v.copy()

    # -- arithmetic ------------------------------------------------------------
    def _need_level(self, ct, n=1):
        if ct.level < n:
            raise RuntimeError("ciphertext level should be positive for multiplication")

    def multiply(self, a, b, relin_key=None):
        if isinstance(a, Ciphertext) and isinstance(b, Ciphertext):
            lvl = min(a.level, b.level)
            if lvl < 1:
                raise RuntimeError("ciphertext level should be positive for multiplication")
            out = self._new(a.v * b.v, lvl - 1)
            self._rec("mul_cc", a.uid, b.uid, out.uid)
            return out
        if isinstance(b, Ciphertext):
            a, b = b, a
        self._need_level(a)
        if isinstance(b, Plaintext):
            out = self._new(a.v * b.v, a.level - 1)
            self._rec("mul_cp", a.uid, b.uid, out.uid)
            return out
        s = complex(b)
        out = self._new(a.v * s, a.level - 1)
        self._rec("mul_cs", a.uid, (round(s.real, 12) + 0.0, round(s.imag, 12) + 0.0), out.uid)
        return out

    def add(self, a, b):
        if isinstance(b, Ciphertext) and not isinstance(a, Ciphertext):
            a, b = b, a
        if isinstance(b, Ciphertext):
            out = self._new(a.v + b.v, min(a.level, b.level))
            self._rec("add_cc", a.uid, b.uid, out.uid)
        elif isinstance(b, Plaintext):
            out = self._new(a.v + b.v, a.level)
            self._rec("add_cp", a.uid, b.uid, out.uid)
        else:
            out = self._new(a.v + complex(b), a.level)
            self._rec("add_cs", a.uid, complex(b), out.uid)
        return out

    def subtract(self, a, b):
        if isinstance(b, Ciphertext):
            out = self._new(a.v - b.v, min(a.level, b.level))
            self._rec("sub_cc", a.uid, b.uid, out.uid)
        elif isinstance(b, Plaintext):
            out = self._new(a.v - b.v, a.level)
            self._rec("sub_cp", a.uid, b.uid, out.uid)
        else:
            out = self._new(a.v - complex(b), a.level)
            self._rec("sub_cs", a.uid, complex(b), out.uid)
        return out

    def add_plain(self, ct, val):
        out = self._new(ct.v + float(val), ct.level)
        self._rec("add_cs", ct.uid, complex(float(val)), out.uid)
        return out

    def make_power_basis(self, ct, degree, relin_key):
        degree = int(degree)
        # depth of ct^degree is ceil(log2(degree)); refuse if it cannot be afforded
        depth = int(np.ceil(np.log2(degree))) if degree > 1 else 0
        if ct.level < depth or (degree > 1 and ct.level < 1):
            raise RuntimeError("ciphertext level should be positive for multiplication")
        out = [ct]
        for k in range(2, degree + 1):
            a, b = out[k // 2 - 1], out[(k + 1) // 2 - 1]
            p = self._new(a.v * b.v, min(a.level, b.level) - 1)
            self._rec("mul_cc", a.uid, b.uid, p.uid)
            out.append(p)
        self._rec("power_basis", ct.uid, degree)
        return out

    def conjugate(self, ct, key):
        out = self._new(np.conj(ct.v), ct.level)
        self._rec("conj", ct.uid, out.uid)
        return out

    def rotate(self, ct, key, steps):
        steps = int(steps)
        out = self._new(np.roll(ct.v, steps), ct.level)
        self._rec("rotate", ct.uid, steps % self.slot_count, out.uid)
        return out

    def relinearize(self, ct, key):
        raise RuntimeError("ciphertext should have 3 polynomials")

    def bootstrap(self, ct, relin, conj, boot):
        out = self._new(ct.v.copy(), self.boot_level)
        self._rec("bootstrap", ct.uid, out.uid)
        return out

    def ntt(self, x):
        self._rec("ntt", getattr(x, "uid", None))
        return x

    def intt(self, x):
        self._rec("intt", getattr(x, "uid", None))
        return x
