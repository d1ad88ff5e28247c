"""CKKS parameter sets -- oracle-side restatement of DESIGN.md spec S1.  TEST INFRASTRUCTURE ONLY.

The engine derives the same chain in C++ (csrc/engine.cu: default_params); tests assert both agree.

The reference never shows its CKKS parameters: `signature=1` selects
desilofhe's built-in bootstrap set, `signature=2` a `max_level` chain.  This module fixes
N = 2^16 (BASELINE.json metric) and derives a prime chain that meets the workload's needs
(SURVEY.md App. B): fresh depth >= 13, post-bootstrap depth >= 5, slot magnitude 256
head-room in q0/Delta, one canonical scale per level.

Chain layout:  q_0 (base, `q0_bits`), q_1..q_L (scale primes near 2^scale_bits), p_0..p_{K-1}
(special primes, `p_bits`).  All primes are = 1 mod 2N.  Scale primes are picked top-down so the
canonical scale S_{l-1} = S_l^2 / q_l stays re-centred on 2^scale_bits (DESIGN.md spec S1).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List


def _is_prime(n: int) -> bool:
    if n < 2:
        return False
    small = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)
    for p in small:
        if n % p == 0:
            return n == p
    d, r = n - 1, 0
    while d % 2 == 0:
        d //= 2
        r += 1
    for a in small:            # deterministic for n < 3.3e24
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(r - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def _primes_below(bound: int, step: int, count: int, used: set) -> List[int]:
    """`count` largest primes p < bound with p = 1 (mod step), skipping `used`."""
    out = []
    c = (bound - 2) // step * step + 1
    while len(out) < count:
        if c not in used and _is_prime(c):
            out.append(c)
            used.add(c)
        c -= step
    return out


def _prime_nearest(target: float, step: int, used: set) -> int:
    """Unused prime = 1 (mod step) nearest to `target` (ties: the smaller one)."""
    k0 = int(round((target - 1) / step))
    for d in range(0, 1 << 20):
        for k in ((k0 - d, k0 + d) if d else (k0,)):
            c = k * step + 1
            if c > 2 and c not in used and _is_prime(c):
                used.add(c)
                return c
    raise RuntimeError("no prime found")


@dataclass
class CKKSParams:
    logn: int
    q: List[int]                  # q_0 .. q_L
    p: List[int]                  # special primes
    scale_bits: int
    alpha: int                    # q-limbs per key-switch digit
    hamming_weight: int
    scales: List[float] = field(default_factory=list)   # canonical scale per level, S[L] = 2^scale_bits
    fresh_level: int = -1         # level of fresh encryptions (<= L)
    boot: dict = field(default_factory=dict)             # bootstrapping plan (see bootstrap.py)
    scale_drop: int = 0           # bits S_0 sits below 2^scale_bits (descending-scale chain; decrypt aligns to level 0)

    @property
    def n(self) -> int:
        return 1 << self.logn

    @property
    def slots(self) -> int:
        return 1 << (self.logn - 1)

    @property
    def L(self) -> int:
        return len(self.q) - 1

    @property
    def K(self) -> int:
        return len(self.p)

    @property
    def dnum(self) -> int:
        return -(-len(self.q) // self.alpha)

    def log_pq(self) -> float:
        import math
        return sum(math.log2(x) for x in self.q + self.p)


RATIO_BITS = 10      # q_0 / S_0 (ModRaise message ratio, head-room of the modulus-256 XOR outputs at level 0)


def make_params(logn: int = 16, levels: int = 20, scale_bits: int = 50, q0_bits: int = 50, p_bits: int = 50,
                alpha: int = 0, dnum: int = 3, hamming_weight: int = 192, fresh_level: int = -1,
                top_levels: int = 0, top_bits: int = 58) -> CKKSParams:
    """Deterministic parameter construction.  `levels` = L (number of scale primes).  With top_levels > 0 the highest
    levels carry the scale 2^top_bits and the chain descends to 2^scale_bits as fast as primes below 2^60.5 allow."""
    step = 2 << logn
    used: set = set()
    nq = levels + 1
    if alpha <= 0:
        alpha = -(-nq // dnum)
    # special primes = the largest primes below 2^p_bits; reserve 16 so `used` protects them, cut to K below
    reserve = _primes_below(1 << p_bits, step, 16, used)
    q0 = _primes_below(1 << q0_bits, step, 1, used)[0]
    hi = float(1 << top_bits)
    cap = math.pow(2.0, 60.5)
    # a q_0 below 2^(scale_bits + 10) keeps the message ratio q_0 / S_0 = 2^10 by letting the scale descend:
    # S_l = 2^(scale_bits - drop 2^-l) (engine.cu `desired_scale`; drop = 0 is the uniform chain)
    drop = max(0, scale_bits + RATIO_BITS - q0_bits)

    def desired(l):
        return math.pow(2.0, scale_bits - drop * math.ldexp(1.0, -l)) if drop > 0 else math.ldexp(1.0, scale_bits)

    scales = [0.0] * nq
    scales[levels] = hi if top_levels > 0 else desired(levels)
    q = [0] * nq
    q[0] = q0
    for l in range(levels, 0, -1):
        delta = desired(l - 1)
        want = hi if (top_levels > 0 and l - 1 > levels - top_levels) else delta       # desired S_{l-1}
        nxt = max(want, scales[l] * scales[l] / cap)
        target = scales[l] * scales[l] / nxt
        q[l] = _prime_nearest(target, step, used)
        scales[l - 1] = scales[l] * scales[l] / float(q[l])
    # P must dominate the widest key-switch digit (alpha consecutive limbs): sum of the limbs' bit lengths
    digit_bits = max(sum(x.bit_length() for x in q[j:j + alpha]) for j in range(0, nq, alpha))
    K = -(-(digit_bits + 1) // (p_bits - 1))
    p = reserve[:K]
    return CKKSParams(logn=logn, q=q, p=p, scale_bits=scale_bits, alpha=alpha, hamming_weight=hamming_weight,
                      scales=scales, fresh_level=levels if fresh_level < 0 else min(fresh_level, levels),
                      scale_drop=drop)


