"""Tier-B oracle: integer RNS-CKKS engine on the CPU.  TEST INFRASTRUCTURE ONLY.

Orchestrates the C primitives of oracle/ckks_ref.c (see that file's header: parity is
UNPINNED at the integer level because the reference's arithmetic is the closed `desilofhe`
wheel; this is a restatement of the published RNS-CKKS construction under DESIGN.md's
"Arithmetic spec").  The CUDA engine must reproduce every integer result of this class
bit-for-bit under identical parameters and seed; floating-point encode/decode follow the
same operation order and are compared bit-exactly where stated, else by tolerance.

Boundary restated: the `desilofhe.Engine` methods used by reference
`engine_context.py:44-204` (encode/encrypt/decrypt/multiply/add/subtract/make_power_basis/
conjugate/rotate/bootstrap), with the semantics of SURVEY.md Appendix A.
Product code must never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB: Optional[C.CDLL] = None

u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def lib() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    so = _HERE / "libckks_ref.so"
    if not so.exists() or so.stat().st_mtime < (_HERE / "ckks_ref.c").stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "libckks_ref.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(str(so))
    u64, i32, sz, dbl = C.c_uint64, C.c_int, C.c_size_t, C.c_double
    sig = {
        "ref_mulmod": (u64, [u64, u64, u64]), "ref_powmod": (u64, [u64, u64, u64]),
        "ref_find_psi": (u64, [u64, i32]), "ref_ntt_tables": (None, [i32, u64, u64, u64p, u64p]),
        "ref_ntt_fwd_batch": (None, [u64p, i32, i32, u64p, u64p]),
        "ref_ntt_inv_batch": (None, [u64p, i32, i32, u64p, u64p]),
        "ref_mul_batch": (None, [u64p, u64p, u64p, i32, sz, u64p]),
        "ref_muladd_batch": (None, [u64p, u64p, u64p, i32, sz, u64p]),
        "ref_add_batch": (None, [u64p, u64p, u64p, i32, sz, u64p]),
        "ref_sub_batch": (None, [u64p, u64p, u64p, i32, sz, u64p]),
        "ref_mul_scalar_batch": (None, [u64p, u64p, u64p, i32, sz, u64p]),
        "ref_mul_const_batch": (None, [u64p, u64p, u64p, u64p, i32, sz, u64p]),
        "ref_permute_batch": (None, [u64p, u64p, u32p, i32, sz]),
        "ref_baseconv": (None, [u64p, u64p, sz, i32, u64p, u64p, i32, u64p, u64p]),
        "ref_baseconv_exact": (None, [u64p, u64p, sz, i32, u64p, u64p, i32, u64p, u64p, f64p, u64p]),
        "ref_reduce_i64_batch": (None, [u64p, i64p, i32, sz, u64p]),
        "ref_rand64": (u64, [u64, u64, u64]),
        "ref_sample_uniform": (None, [u64p, sz, u64, u64, u64, u64]),
        "ref_sample_cbd": (None, [i64p, sz, u64, u64]), "ref_sample_ternary": (None, [i64p, sz, u64, u64]),
        "ref_sample_sparse": (None, [i64p, sz, i32, u64, u64]),
        "ref_fft_tables": (None, [i32, u32p, f64p]),
        "ref_special_fft": (None, [f64p, i32, u32p, f64p]), "ref_special_ifft": (None, [f64p, i32, u32p, f64p]),
        "ref_round_coeffs": (i32, [i64p, f64p, sz, dbl]),
        "ref_center_to_w": (None, [f64p, u64p, sz, u64, dbl]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _LIB = L
    return L


# stream ids of the sampling spec: kind << 48 | a << 16 | b
def stream_id(kind: int, a: int = 0, b: int = 0) -> int:
    return (kind << 48) | (a << 16) | b


ST_SK, ST_PK_A, ST_PK_E, ST_EVK_A, ST_EVK_E, ST_ENC_V, ST_ENC_E0, ST_ENC_E1 = range(1, 9)
RELIN_ID = 0   # "galois element" slot used for the relinearisation key in evk stream ids


def bitrev(x: int, bits: int) -> int:
    return int(format(x, f"0{bits}b")[::-1], 2) if bits else 0


class Ct:
    """Ciphertext: polys [npoly][level+1][N] uint64 (NTT domain, bit-reversed order)."""
    __slots__ = ("c", "level", "scale")

    def __init__(self, c: np.ndarray, level: int, scale: float):
        self.c, self.level, self.scale = c, level, scale


class OracleCKKS:
    def __init__(self, params, seed: int = 1):
        self.P = params
        self.L = params.L
        self.logn, self.N, self.n = params.logn, params.n, params.slots
        self.q: List[int] = list(params.q)
        self.p: List[int] = list(params.p)
        self.K = len(self.p)
        self.alpha = params.alpha
        self.moduli = self.q + self.p                       # global limb index: q_i -> i, p_k -> L+1+k
        self.seed = int(seed)
        self.scales = list(params.scales)
        self.fresh_level = params.fresh_level
        self.lib = lib()
        nm = len(self.moduli)
        self.mod_arr = np.array(self.moduli, dtype=np.uint64)
        self.psi = [self.lib.ref_find_psi(m, self.logn) for m in self.moduli]
        self.tab = np.zeros((nm, self.N), dtype=np.uint64)
        self.itab = np.zeros((nm, self.N), dtype=np.uint64)
        for i, m in enumerate(self.moduli):
            self.lib.ref_ntt_tables(self.logn, m, self.psi[i], self.tab[i], self.itab[i])
        # J = psi^(N/2): the NTT-domain image of X^(N/2) is +-J (spec S7)
        self.J = [pow(self.psi[i], self.N // 2, m) for i, m in enumerate(self.moduli)]
        self.rot = np.zeros(self.n, dtype=np.uint32)
        self.ksi = np.zeros(2 * (2 * self.N + 1), dtype=np.float64)
        self.lib.ref_fft_tables(self.logn, self.rot, self.ksi)
        self._perm_cache: Dict[int, np.ndarray] = {}
        self._enc_counter = 0
        self.sk_coef: Optional[np.ndarray] = None
        self.sk_ntt: Optional[np.ndarray] = None            # [L+1+K][N]
        self.pk: Optional[np.ndarray] = None
        self.evk: Dict[int, np.ndarray] = {}                 # key id -> [dnum][2][L+1+K][N]
        self.counters: Dict[str, int] = {}

    # ------------------------------------------------------------------ limb-set helpers
    def _idx_q(self, level: int) -> List[int]:
        return list(range(level + 1))

    def _idx_qp(self, level: int) -> List[int]:
        return list(range(level + 1)) + list(range(self.L + 1, self.L + 1 + self.K))

    def _mods(self, idx: Sequence[int]) -> np.ndarray:
        return np.ascontiguousarray(self.mod_arr[list(idx)])

    def ntt(self, a: np.ndarray, idx: Sequence[int]) -> np.ndarray:
        a = np.ascontiguousarray(a.copy())
        self.lib.ref_ntt_fwd_batch(a, len(idx), self.logn, self._mods(idx), np.ascontiguousarray(self.tab[list(idx)]))
        self.counters["ntt"] = self.counters.get("ntt", 0) + len(idx)
        return a

    def intt(self, a: np.ndarray, idx: Sequence[int]) -> np.ndarray:
        a = np.ascontiguousarray(a.copy())
        self.lib.ref_ntt_inv_batch(a, len(idx), self.logn, self._mods(idx), np.ascontiguousarray(self.itab[list(idx)]))
        self.counters["ntt"] = self.counters.get("ntt", 0) + len(idx)
        return a

    def _binop(self, fn, a, b, idx):
        out = np.empty_like(a)
        fn(out, np.ascontiguousarray(a), np.ascontiguousarray(b), len(idx), self.N, self._mods(idx))
        return out

    def mul(self, a, b, idx):
        return self._binop(self.lib.ref_mul_batch, a, b, idx)

    def add(self, a, b, idx):
        return self._binop(self.lib.ref_add_batch, a, b, idx)

    def sub(self, a, b, idx):
        return self._binop(self.lib.ref_sub_batch, a, b, idx)

    def mul_scalar(self, a, scal: Sequence[int], idx):
        out = np.empty_like(a)
        s = np.array([int(x) % self.moduli[i] for x, i in zip(scal, idx)], dtype=np.uint64)
        self.lib.ref_mul_scalar_batch(out, np.ascontiguousarray(a), s, len(idx), self.N, self._mods(idx))
        return out

    def reduce_i64(self, v: np.ndarray, idx) -> np.ndarray:
        out = np.empty((len(idx), self.N), dtype=np.uint64)
        self.lib.ref_reduce_i64_batch(out, np.ascontiguousarray(v, dtype=np.int64), len(idx), self.N, self._mods(idx))
        return out

    # ------------------------------------------------------------------ Galois maps (spec S4)
    def galois_for_rotation(self, steps: int) -> int:
        """rotate(ct, +r) == np.roll(slots, +r)  <=>  X -> X^(5^(-r mod n))   (SURVEY.md App. E)."""
        r = steps % self.n
        return pow(5, (self.n - r) % self.n, 2 * self.N)

    def galois_conj(self) -> int:
        return 2 * self.N - 1

    def galois_perm(self, g: int) -> np.ndarray:
        """NTT-domain gather: out[k] = in[perm[k]] realises m(X) -> m(X^g)."""
        if g not in self._perm_cache:
            M, bits = 2 * self.N, self.logn
            k = np.arange(self.N, dtype=np.int64)
            br = np.array([bitrev(int(x), bits) for x in range(self.N)], dtype=np.int64)
            e = (2 * br[k] + 1) * g % M
            self._perm_cache[g] = br[(e - 1) // 2].astype(np.uint32)
        return self._perm_cache[g]

    def automorph(self, a: np.ndarray, g: int) -> np.ndarray:
        out = np.empty_like(a)
        self.lib.ref_permute_batch(out, np.ascontiguousarray(a), self.galois_perm(g), a.shape[0], self.N)
        return out

    # ------------------------------------------------------------------ sampling / keys (spec S8)
    def _uniform(self, idx, stream) -> np.ndarray:
        out = np.empty((len(idx), self.N), dtype=np.uint64)
        for r, i in enumerate(idx):
            self.lib.ref_sample_uniform(out[r], self.N, self.moduli[i], self.seed, stream, i)
        return out

    def _cbd(self, stream) -> np.ndarray:
        e = np.empty(self.N, dtype=np.int64)
        self.lib.ref_sample_cbd(e, self.N, self.seed, stream)
        return e

    def keygen_secret(self):
        s = np.empty(self.N, dtype=np.int64)
        self.lib.ref_sample_sparse(s, self.N, self.P.hamming_weight, self.seed, stream_id(ST_SK))
        self.sk_coef = s
        allidx = list(range(len(self.moduli)))
        self.sk_ntt = self.ntt(self.reduce_i64(s, allidx), allidx)
        return s

    def keygen_public(self):
        idx = self._idx_q(self.L)
        a = self._uniform(idx, stream_id(ST_PK_A))
        e = self.ntt(self.reduce_i64(self._cbd(stream_id(ST_PK_E)), idx), idx)
        b = self.sub(e, self.mul(a, self.sk_ntt[idx], idx), idx)
        self.pk = np.stack([b, a])
        return self.pk

    def _digit_limbs(self, j: int, level: int) -> List[int]:
        return [i for i in range(j * self.alpha, min((j + 1) * self.alpha, level + 1))]

    def keygen_switch(self, key_id: int, s_from_ntt: np.ndarray) -> np.ndarray:
        """evk[j] = (-a_j s + e_j + P * F_j * s_from,  a_j) over Q_L u P; F_j = CRT selector of digit j."""
        idx = self._idx_qp(self.L)
        Pprod = 1
        for x in self.p:
            Pprod *= x
        dn = self.P.dnum
        out = np.zeros((dn, 2, len(idx), self.N), dtype=np.uint64)
        for j in range(dn):
            a = self._uniform(idx, stream_id(ST_EVK_A, key_id, j))
            e = self.ntt(self.reduce_i64(self._cbd(stream_id(ST_EVK_E, key_id, j)), idx), idx)
            b = self.sub(e, self.mul(a, self.sk_ntt[idx], idx), idx)
            dl = self._digit_limbs(j, self.L)
            fac = [(Pprod % self.moduli[i]) if i in dl else 0 for i in idx]
            b = self.add(b, self.mul_scalar(s_from_ntt[idx], fac, idx), idx)
            out[j, 0], out[j, 1] = b, a
        self.evk[key_id] = out
        return out

    def keygen_relin(self):
        allidx = list(range(len(self.moduli)))
        s2 = self.mul(self.sk_ntt, self.sk_ntt, allidx)
        return self.keygen_switch(RELIN_ID, s2)

    def keygen_galois(self, g: int):
        return self.keygen_switch(g, self.automorph(self.sk_ntt, g))

    # ------------------------------------------------------------------ encode / decode (spec S9)
    def const_residues(self, c: complex, scale: float, idx) -> Tuple[np.ndarray, np.ndarray]:
        """Residues of the constant polynomial R + I X^(N/2): (R + I J, R - I J) per limb."""
        R, I = int(np.rint(c.real * scale)), int(np.rint(c.imag * scale))
        cp = np.array([(R + I * self.J[i]) % self.moduli[i] for i in idx], dtype=np.uint64)
        cm = np.array([(R - I * self.J[i]) % self.moduli[i] for i in idx], dtype=np.uint64)
        return cp, cm

    def encode_coeffs(self, z: np.ndarray, scale: float) -> np.ndarray:
        """slots -> N signed integer coefficients at `scale`."""
        v = np.ascontiguousarray(np.asarray(z, dtype=np.complex128)).view(np.float64).copy()
        self.lib.ref_special_ifft(v, self.logn, self.rot, self.ksi)
        out = np.empty(self.N, dtype=np.int64)
        if self.lib.ref_round_coeffs(out, v, self.n, float(scale)):
            raise OverflowError("plaintext coefficient does not fit 62 bits")
        return out

    def encode(self, z: np.ndarray, level: int, scale: float) -> np.ndarray:
        idx = self._idx_q(level)
        return self.ntt(self.reduce_i64(self.encode_coeffs(z, scale), idx), idx)

    def decode_limb0(self, coef: np.ndarray, scale: float) -> np.ndarray:
        w = np.empty(2 * self.n, dtype=np.float64)
        self.lib.ref_center_to_w(w, np.ascontiguousarray(coef), self.n, self.q[0], float(scale))
        self.lib.ref_special_fft(w, self.logn, self.rot, self.ksi)
        return w.view(np.complex128).copy()

    # ------------------------------------------------------------------ encrypt / decrypt
    def encrypt(self, z: np.ndarray, level: Optional[int] = None) -> Ct:
        level = self.fresh_level if level is None else level
        scale = self.scales[level]
        idx = self._idx_q(level)
        m = self.encode(z, level, scale)
        k = self._enc_counter
        self._enc_counter += 1
        v = np.empty(self.N, dtype=np.int64)
        self.lib.ref_sample_ternary(v, self.N, self.seed, stream_id(ST_ENC_V, k))
        v = self.ntt(self.reduce_i64(v, idx), idx)
        e0 = self.ntt(self.reduce_i64(self._cbd(stream_id(ST_ENC_E0, k)), idx), idx)
        e1 = self.ntt(self.reduce_i64(self._cbd(stream_id(ST_ENC_E1, k)), idx), idx)
        c0 = self.add(self.add(self.mul(v, self.pk[0][idx], idx), e0, idx), m, idx)
        c1 = self.add(self.mul(v, self.pk[1][idx], idx), e1, idx)
        return Ct(np.stack([c0, c1]), level, scale)

    def decrypt(self, ct: Ct) -> np.ndarray:
        """Limb 0 only: |m*scale + e| < q0/2 at every level (spec S10); a descending-scale chain (q_0 next to the scale
        primes) aligns to level 0 first, where the scale is 2^10 below q_0."""
        if getattr(self.P, "scale_drop", 0) > 0 and ct.level > 0:
            ct = self.level_down(ct, 0)
        idx = [0]
        t = ct.c[0][:1]
        spow = self.sk_ntt[:1]
        for k in range(1, ct.c.shape[0]):
            t = self.add(t, self.mul(ct.c[k][:1], spow, idx), idx)
            spow = self.mul(spow, self.sk_ntt[:1], idx)
        coef = self.intt(t, idx)[0]
        return self.decode_limb0(coef, ct.scale)

    # ------------------------------------------------------------------ key switching (spec S5/S6)
    def _baseconv(self, x: np.ndarray, src: List[int], tgt: List[int], exact: bool = False) -> np.ndarray:
        D = 1
        for i in src:
            D *= self.moduli[i]
        hatinv = np.array([pow(D // self.moduli[i] % self.moduli[i], -1, self.moduli[i]) for i in src], dtype=np.uint64)
        hat = np.array([[D // self.moduli[i] % self.moduli[t] for t in tgt] for i in src], dtype=np.uint64)
        out = np.empty((len(tgt), self.N), dtype=np.uint64)
        if exact:      # spec S5': centred residue modulo D (ModDown)
            inv_src = np.array([1.0 / float(self.moduli[i]) for i in src], dtype=np.float64)
            negD = np.array([(self.moduli[t] - D % self.moduli[t]) % self.moduli[t] for t in tgt], dtype=np.uint64)
            self.lib.ref_baseconv_exact(out, np.ascontiguousarray(x), self.N, len(src), self._mods(src), hatinv,
                                        len(tgt), self._mods(tgt), np.ascontiguousarray(hat), inv_src, negD)
            return out
        self.lib.ref_baseconv(out, np.ascontiguousarray(x), self.N, len(src), self._mods(src), hatinv, len(tgt),
                              self._mods(tgt), np.ascontiguousarray(hat))
        return out

    def mod_up(self, d: np.ndarray, level: int) -> List[np.ndarray]:
        """NTT-domain poly [level+1][N] -> per digit its extension over Q_level u P (NTT domain)."""
        qp = self._idx_qp(level)
        beta = -(-(level + 1) // self.alpha)
        out = []
        for j in range(beta):
            dl = self._digit_limbs(j, level)
            coef = self.intt(d[dl], dl)
            others = [i for i in qp if i not in dl]
            ext = self.ntt(self._baseconv(coef, dl, others), others)
            full = np.empty((len(qp), self.N), dtype=np.uint64)
            for r, i in enumerate(qp):
                full[r] = d[i] if i in dl else ext[others.index(i)]
            out.append(full)
        return out

    def mod_down(self, acc: np.ndarray, level: int, drop: int = 0) -> np.ndarray:
        """[level+1+K][N] over Q_level u P  ->  floor(acc / (P q_{level-drop+1} .. q_level)) over Q_{level-drop}
        (NTT domain).  drop = 0 is the plain ModDown; drop = 1, 2 merge one or two rescales into it (spec S6b)."""
        nq = level + 1
        keep = self._idx_q(level - drop)
        src_mods = list(range(level - drop + 1, level + 1)) + list(range(self.L + 1, self.L + 1 + self.K))
        src_rows = list(range(level - drop + 1, level + 1)) + list(range(nq, nq + self.K))
        coef = self.intt(acc[src_rows], src_mods)
        conv = self.ntt(self._baseconv(coef, src_mods, keep, exact=True), keep)
        D = 1
        for i in src_mods:
            D *= self.moduli[i]
        inv = [pow(D % self.moduli[i], -1, self.moduli[i]) for i in keep]
        return self.mul_scalar(self.sub(acc[:len(keep)], conv, keep), inv, keep)

    def key_switch(self, d: np.ndarray, level: int, key_id: int, digits: Optional[List[np.ndarray]] = None,
                   addend: Optional[np.ndarray] = None, drop: int = 0):
        """addend = (d0, d1) over Q_level is folded in as P * addend before the division (merged relinearisation)."""
        self.counters["keyswitch"] = self.counters.get("keyswitch", 0) + 1
        qp = self._idx_qp(level)
        rows = list(range(level + 1)) + list(range(self.L + 1, self.L + 1 + self.K))   # rows into evk's limb axis
        digits = self.mod_up(d, level) if digits is None else digits
        evk = self.evk[key_id]
        acc0 = np.zeros((len(qp), self.N), dtype=np.uint64)
        acc1 = np.zeros((len(qp), self.N), dtype=np.uint64)
        mods = self._mods(qp)
        for j, ext in enumerate(digits):
            self.lib.ref_muladd_batch(acc0, ext, np.ascontiguousarray(evk[j, 0][rows]), len(qp), self.N, mods)
            self.lib.ref_muladd_batch(acc1, ext, np.ascontiguousarray(evk[j, 1][rows]), len(qp), self.N, mods)
        if addend is not None:
            qi = self._idx_q(level)
            Pprod = 1
            for x in self.p:
                Pprod *= x
            pm = [Pprod % self.moduli[i] for i in qi]
            acc0[:level + 1] = self.add(acc0[:level + 1], self.mul_scalar(addend[0], pm, qi), qi)
            acc1[:level + 1] = self.add(acc1[:level + 1], self.mul_scalar(addend[1], pm, qi), qi)
        return self.mod_down(acc0, level, drop), self.mod_down(acc1, level, drop)

    # ------------------------------------------------------------------ rescale / level management (spec S6)
    def rescale_poly(self, c: np.ndarray, level: int) -> np.ndarray:
        ql = self.q[level]
        h = ql // 2
        last = self.intt(c[level:level + 1], [level])[0]
        t = (last + np.uint64(h)) % np.uint64(ql)
        lo = self._idx_q(level - 1)
        delta = np.empty((level, self.N), dtype=np.uint64)
        for i in lo:
            qi = np.uint64(self.q[i])
            r = t % qi
            hm = np.uint64(h % self.q[i])
            delta[i] = np.where(r >= hm, r - hm, r + qi - hm)
        delta = self.ntt(delta, lo)
        inv = [pow(ql % self.q[i], -1, self.q[i]) for i in lo]
        return self.mul_scalar(self.sub(c[:level], delta, lo), inv, lo)

    def rescale(self, ct: Ct) -> Ct:
        if ct.level < 1:
            raise RuntimeError("ciphertext level should be positive for multiplication")
        c = np.stack([self.rescale_poly(ct.c[k], ct.level) for k in range(ct.c.shape[0])])
        return Ct(c, ct.level - 1, ct.scale / float(self.q[ct.level]))

    def level_down(self, ct: Ct, target: int) -> Ct:
        """Canonical-scale alignment: drop to target+1, multiply by round(S_t q_{t+1} / S_l), rescale."""
        if target == ct.level:
            return ct
        assert target < ct.level
        t1 = target + 1
        idx = self._idx_q(t1)
        k = int(np.rint(self.scales[target] * float(self.q[t1]) / ct.scale))
        c = np.stack([self.mul_scalar(ct.c[j][:t1 + 1], [k] * len(idx), idx) for j in range(ct.c.shape[0])])
        out = self.rescale(Ct(c, t1, ct.scale * k))
        out.scale = self.scales[target]
        return out

    def align(self, a: Ct, b: Ct) -> Tuple[Ct, Ct]:
        l = min(a.level, b.level)
        return self.level_down(a, l), self.level_down(b, l)

    # ------------------------------------------------------------------ homomorphic ops
    def add_ct(self, a: Ct, b: Ct) -> Ct:
        a, b = self.align(a, b)
        idx = self._idx_q(a.level)
        return Ct(np.stack([self.add(a.c[k], b.c[k], idx) for k in range(2)]), a.level, a.scale)

    def sub_ct(self, a: Ct, b: Ct) -> Ct:
        a, b = self.align(a, b)
        idx = self._idx_q(a.level)
        return Ct(np.stack([self.sub(a.c[k], b.c[k], idx) for k in range(2)]), a.level, a.scale)

    def pt_scale(self, level: int) -> float:
        """Scale at which a plaintext meets a level-`level` ciphertext so the product lands on S[level-1]: the canonical
        scale S[level] itself (DESIGN.md spec S1/S7: S[l-1] = S[l]^2 / q_l).  The algebraically equal expression
        q_l * S[l-1] / S[l] differs from S[l] by one ulp at levels 5, 8, 11, 14, 16 of the benchmark's 22-prime chain --
        enough to change round(c * scale) for a large constant; found by the bench-chain parity case
        (tests/test_engine_parity.py, cuda-n16-bench) and resolved in favour of the written spec."""
        return self.scales[level]

    def mul_const(self, a: Ct, c: complex) -> Ct:
        if a.level < 1:
            raise RuntimeError("ciphertext level should be positive for multiplication")
        idx = self._idx_q(a.level)
        s = self.pt_scale(a.level)
        cp, cm = self.const_residues(complex(c), s, idx)
        out = np.empty_like(a.c)
        for k in range(2):
            self.lib.ref_mul_const_batch(out[k], np.ascontiguousarray(a.c[k]), cp, cm, len(idx), self.N, self._mods(idx))
        r = self.rescale(Ct(out, a.level, a.scale * s))
        r.scale = self.scales[r.level]
        return r

    def mul_plain_vec(self, a: Ct, z: np.ndarray) -> Ct:
        if a.level < 1:
            raise RuntimeError("ciphertext level should be positive for multiplication")
        idx = self._idx_q(a.level)
        s = self.pt_scale(a.level)
        m = self.encode(z, a.level, s)
        r = self.rescale(Ct(np.stack([self.mul(a.c[k], m, idx) for k in range(2)]), a.level, a.scale * s))
        r.scale = self.scales[r.level]
        return r

    def add_const(self, a: Ct, c: complex) -> Ct:
        idx = self._idx_q(a.level)
        cp, cm = self.const_residues(complex(c), a.scale, idx)
        c0 = a.c[0].copy()
        h = self.N // 2
        for r, i in enumerate(idx):
            qi = np.uint64(self.moduli[i])
            c0[r, :h] = (c0[r, :h] + cp[r]) % qi
            c0[r, h:] = (c0[r, h:] + cm[r]) % qi
        return Ct(np.stack([c0, a.c[1]]), a.level, a.scale)

    def add_plain_vec(self, a: Ct, z: np.ndarray) -> Ct:
        idx = self._idx_q(a.level)
        m = self.encode(z, a.level, a.scale)
        return Ct(np.stack([self.add(a.c[0], m, idx), a.c[1]]), a.level, a.scale)

    def tensor(self, a: Ct, b: Ct) -> Ct:
        a, b = self.align(a, b)
        idx = self._idx_q(a.level)
        d0 = self.mul(a.c[0], b.c[0], idx)
        d1 = self.add(self.mul(a.c[0], b.c[1], idx), self.mul(a.c[1], b.c[0], idx), idx)
        d2 = self.mul(a.c[1], b.c[1], idx)
        return Ct(np.stack([d0, d1, d2]), a.level, a.scale * b.scale)

    def relinearize(self, t: Ct) -> Ct:
        idx = self._idx_q(t.level)
        k0, k1 = self.key_switch(t.c[2], t.level, RELIN_ID)
        return Ct(np.stack([self.add(t.c[0], k0, idx), self.add(t.c[1], k1, idx)]), t.level, t.scale)

    def mul_ct(self, a: Ct, b: Ct) -> Ct:
        if min(a.level, b.level) < 1:
            raise RuntimeError("ciphertext level should be positive for multiplication")
        self.counters["mul_cc"] = self.counters.get("mul_cc", 0) + 1
        t = self.tensor(a, b)
        k0, k1 = self.key_switch(t.c[2], t.level, RELIN_ID, addend=t.c[:2], drop=1)      # spec S6b
        return Ct(np.stack([k0, k1]), t.level - 1, self.scales[t.level - 1])

    def apply_galois(self, a: Ct, g: int) -> Ct:
        if g not in self.evk:
            self.keygen_galois(g)
        idx = self._idx_q(a.level)
        c0 = self.automorph(a.c[0], g)
        c1 = self.automorph(a.c[1], g)
        k0, k1 = self.key_switch(c1, a.level, g)
        return Ct(np.stack([self.add(c0, k0, idx), k1]), a.level, a.scale)

    def rotate(self, a: Ct, steps: int) -> Ct:
        if steps % self.n == 0:
            return a
        return self.apply_galois(a, self.galois_for_rotation(steps))

    def conjugate(self, a: Ct) -> Ct:
        return self.apply_galois(a, self.galois_conj())

    def power_basis(self, a: Ct, degree: int) -> List[Ct]:
        """[a^1 .. a^degree], a^k = a^(k//2) * a^((k+1)//2)  (depth ceil(log2 k))."""
        depth = int(np.ceil(np.log2(degree))) if degree > 1 else 0
        if a.level < depth:
            raise RuntimeError("ciphertext level should be positive for multiplication")
        out = [a]
        for k in range(2, degree + 1):
            out.append(self.mul_ct(out[k // 2 - 1], out[(k + 1) // 2 - 1]))
        return out

    # ------------------------------------------------------------------ fused LUT evaluation (csrc/lut.cu)
    def _mul_const_raw(self, poly: np.ndarray, c: complex, scale: float, idx) -> np.ndarray:
        cp, cm = self.const_residues(complex(c), scale, idx)
        out = np.empty_like(poly)
        self.lib.ref_mul_const_batch(out, np.ascontiguousarray(poly), cp, cm, len(idx), self.N, self._mods(idx))
        return out

    def lut2(self, A: Dict[int, Ct], B: Dict[int, Ct], terms) -> Ct:
        """sum_t c_t A[p_t] (x) B[q_t]: terms grouped by p, constants at scale S[l-1], ONE relinearisation,
        two rescales -> level l-2 (same spec as Engine::lut2)."""
        level = min(min(A[p].level, B[q].level) for p, q, _ in terms)
        if level < 2:
            raise RuntimeError("ciphertext level should be positive for multiplication")
        idx = self._idx_q(level)
        order = sorted(range(len(terms)), key=lambda t: terms[t][0])        # stable
        d = np.zeros((3, level + 1, self.N), dtype=np.uint64)
        mods = self._mods(idx)
        k = 0
        while k < len(order):
            p = terms[order[k]][0]
            u0 = np.zeros((level + 1, self.N), dtype=np.uint64)
            u1 = np.zeros((level + 1, self.N), dtype=np.uint64)
            while k < len(order) and terms[order[k]][0] == p:
                _, q, c = terms[order[k]]
                b = self.level_down(B[q], level)
                u0 = self.add(u0, self._mul_const_raw(b.c[0], c, self.scales[level - 1], idx), idx)
                u1 = self.add(u1, self._mul_const_raw(b.c[1], c, self.scales[level - 1], idx), idx)
                k += 1
            a = self.level_down(A[p], level)
            a0, a1 = np.ascontiguousarray(a.c[0]), np.ascontiguousarray(a.c[1])
            self.lib.ref_muladd_batch(d[0], a0, u0, len(idx), self.N, mods)
            self.lib.ref_muladd_batch(d[1], a0, u1, len(idx), self.N, mods)
            self.lib.ref_muladd_batch(d[1], a1, u0, len(idx), self.N, mods)
            self.lib.ref_muladd_batch(d[2], a1, u1, len(idx), self.N, mods)
        self.counters["mul_cc"] = self.counters.get("mul_cc", 0) + 1
        k0, k1 = self.key_switch(d[2], level, RELIN_ID, addend=d[:2], drop=2)             # spec S6b
        return Ct(np.stack([k0, k1]), level - 2, self.scales[level - 2])

    def lincomb(self, X: Sequence[Ct], coeffs) -> Ct:
        """sum_k c_k X_k: per distinct level one un-rescaled multiply-accumulate and one rescale; partial sums are
        added from the highest level down (same spec as Engine::lincomb)."""
        by_level: Dict[int, List[int]] = {}
        for k, x in enumerate(X):
            if x.level < 1:
                raise RuntimeError("ciphertext level should be positive for multiplication")
            by_level.setdefault(x.level, []).append(k)
        acc = None
        for level in sorted(by_level, reverse=True):
            idx = self._idx_q(level)
            npoly = X[by_level[level][0]].c.shape[0]
            s = np.zeros((npoly, level + 1, self.N), dtype=np.uint64)
            for k in by_level[level]:
                for j in range(npoly):
                    s[j] = self.add(s[j], self._mul_const_raw(X[k].c[j], complex(coeffs[k]), self.scales[level], idx), idx)
            r = self.rescale(Ct(s, level, self.scales[level] ** 2))
            r.scale = self.scales[r.level]
            acc = r if acc is None else self.add_ct(acc, r)
        return acc

    def rotate_hoisted(self, a: Ct, steps: Sequence[int]) -> List[Ct]:
        """Rotations sharing one ModUp: the Galois gather is applied to the decomposed digits (csrc: rotate_hoisted)."""
        idx = self._idx_q(a.level)
        digits = None
        out = []
        for s in steps:
            if s % self.n == 0:
                out.append(a)
                continue
            if digits is None:
                digits = self.mod_up(a.c[1], a.level)
            g = self.galois_for_rotation(s)
            if g not in self.evk:
                self.keygen_galois(g)
            k0, k1 = self.key_switch(None, a.level, g, digits=[self.automorph(d, g) for d in digits])
            out.append(Ct(np.stack([self.add(self.automorph(a.c[0], g), k0, idx), k1]), a.level, a.scale))
        return out
