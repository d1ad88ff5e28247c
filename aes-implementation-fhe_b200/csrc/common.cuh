// common.cuh -- 64-bit modular arithmetic shared by every kernel of the CKKS engine.
//
// All residues are canonical uint64 in [0,q), q < 2^62 (so 4q fits a word and the Harvey lazy
// butterflies of ntt.cu are safe).  Two multiplication flavours:
//   * shoup_mul:   x * w mod q for a *precomputed* w with companion w' = floor(w 2^64 / q)
//                  (1 mul.hi + 2 mul.lo), used for twiddles and per-limb scalars;
//   * barrett_mul: a * b mod q for two variable operands, single-word Barrett on the
//                  128-bit product with a per-modulus shift (2 mul.hi + 2 mul.lo).
// The per-modulus constants live in a small global-memory table (ModConst[], L1/const-cache
// resident; every CTA works on one limb so the lookups are warp-uniform).
#pragma once
#include "platform.cuh"

#define CKKS_MAX_MODULI 64

// per-modulus constants, indexed by global modulus index (q_0..q_L, then p_0..p_{K-1})
struct ModConst {
    u64 q;
    u64 mu;      // floor(2^(k+63)/q), k = bitlen(q): mulhi(z >> (k-1), mu) = floor(z/q) - {0,1,2}
    u32 k1;      // k - 1
    u32 pad;
    u64 ninv;    // N^-1 mod q
    u64 ninv_s;  // Shoup companion of ninv
    u64 w1n;     // psi^-bitrev(1) * N^-1 (last inverse-NTT stage twiddle with the 1/N folded in)
    u64 w1n_s;
};

// list of modulus indices a batched kernel works on: row r of a [rows][N] array is a residue
// polynomial modulo mc[idx[r]].  Passed by value (no H2D copy per launch).
struct LimbList {
    int n;
    unsigned char idx[CKKS_MAX_MODULI];
};
// per-row scalars with Shoup companions, passed by value
struct ScalarList {
    u64 v[CKKS_MAX_MODULI];
    u64 vs[CKKS_MAX_MODULI];
};

__device__ __forceinline__ u64 add_mod(u64 a, u64 b, u64 q) {
    u64 s = a + b;
    return s >= q ? s - q : s;
}
__device__ __forceinline__ u64 sub_mod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
__device__ __forceinline__ u64 neg_mod(u64 a, u64 q) { return a ? q - a : 0; }

// x * w mod q, result in [0, 2q)   (x arbitrary 64-bit, w < q)
__device__ __forceinline__ u64 shoup_mul_lazy(u64 x, u64 w, u64 ws, u64 q) {
    u64 h = mulhi64(ws, x);
    return w * x - h * q;
}
__device__ __forceinline__ u64 shoup_mul(u64 x, u64 w, u64 ws, u64 q) {
    u64 r = shoup_mul_lazy(x, w, ws, q);
    return r >= q ? r - q : r;
}

// reduce the 128-bit value z = (hi,lo) with z >> (k-1) < 2^64 to [0,q); up to `8q` of slack is
// folded by conditional subtractions so sums of a few products (< 16 q^2) are accepted.
__device__ __forceinline__ u64 barrett_reduce128(u64 hi, u64 lo, const ModConst& m) {
    u64 x = (lo >> m.k1) | (hi << (64 - m.k1));       // floor(z / 2^(k-1)); k1 in [1,63]
    u64 qh = mulhi64(x, m.mu);
    u64 r = lo - qh * m.q;
    if (r >= 8 * m.q) r -= 8 * m.q;
    if (r >= 4 * m.q) r -= 4 * m.q;
    if (r >= 2 * m.q) r -= 2 * m.q;
    if (r >= m.q) r -= m.q;
    return r;
}
__device__ __forceinline__ u64 barrett_mul(u64 a, u64 b, const ModConst& m) {
    return barrett_reduce128(mulhi64(a, b), a * b, m);
}
// reduce one word x < 2^64 modulo q (via the same Barrett constant)
__device__ __forceinline__ u64 barrett_reduce64(u64 x, const ModConst& m) { return barrett_reduce128(0, x, m); }

__device__ __forceinline__ void mac128(u64& hi, u64& lo, u64 a, u64 b) {
    u64 pl = a * b, ph = mulhi64(a, b);
    lo += pl;
    hi += ph + (lo < pl);
}

// ------------------------------------------------------------------ FP64-pipe modular arithmetic (moduli below CKKS_FP_LIMIT)
// The B200's full-rate FP64 pipe multiplies modulo a ~2^50 prime about three times as fast as the 64-bit integer
// multiplier (profiles/r1_pipe_peaks.txt, profiles/r2_fp64_round.txt): the NTT butterflies of the scale primes (ntt.cu) and
// the basis conversion towards them (kernels.cu: k_base_convert_fp) hold residues as exact integers in doubles.
#define CKKS_FP_LIMIT 1238489897526886ull     /* 1.1 * 2^50: lazy values up to 5.7 q stay below 2^53 = 7.27 q (tools/ntt_fp_bounds.py) */
// a * w mod q as an exact integer: a any integer with |a| < 2^53, |w| < q < CKKS_FP_LIMIT, qinv = fl(1 / q).
//   h = fl(a w), l = a w - h (exact, FMA), c = rint(h qinv) (quotient, three roundings off), r = (h - c q) + l (both exact)
//   |r| <= (0.5 + 1.5 |a| 2^-52) q.  No companion word per twiddle: the quotient estimate comes from the product itself
//   (half the table traffic of a Shoup-style w / q word, the same five FP64-pipe instructions and one rounding).
__device__ __forceinline__ double modmul_fp(double a, double w, double q, double qinv) {
    const double h = fmul_rn(a, w);
    const double l = ffma_rn(a, w, -h);
    const double c = frint(fmul_rn(h, qinv));
    return fadd_rn(ffma_rn(-c, q, h), l);
}
// fold a lazy value (|x| < 2^53) to |x| <= q/2 (+ one q when the quotient estimate is off by one)
__device__ __forceinline__ double fold_fp(double x, double q, double qinv) {
    return ffma_rn(-frint(fmul_rn(x, qinv)), q, x);
}
// exact canonical residue in [0,q) as an integer
// The fold leaves |x| <= (1/2 + 2^-49) q (the quotient estimate x * qinv is within 2^-49 of x / q for |x| < 2^53, so it can
// only round to the other neighbour next to a tie): ONE correction, +q for a negative value, lands in [0, q), and it is
// done on the integer pipe after the conversion (the FP64 pipe, which bounds these kernels, keeps its three instructions).
__device__ __forceinline__ u64 canon_fp(double x, double q, double qinv) {
    const i64 r = d2ll_rn(fold_fp(x, q, qinv));
    return (u64)(r < 0 ? r + (i64)d2ll_rn(q) : r);
}
