/* oracle/ckks_ref.c -- CPU restatement of the RNS-CKKS arithmetic.  TEST INFRASTRUCTURE ONLY.
 *
 * parity: UNPINNED at the ciphertext/integer level.  The reference's arithmetic lives in the
 * closed, unpinned `desilofhe` wheel (only mention: reference engine_context.py:1); the reference
 * ships no golden ciphertexts, NTT/key-switch KATs or RNG seeds.  This file
 * therefore restates the published full-RNS CKKS construction (Cheon-Han-Kim-Kim-Song SAC'18 RNS
 * variant; hybrid key switching of Han-Ki CT-RSA'20; HEAAN's "special FFT" canonical embedding)
 * under the written spec in DESIGN.md "Arithmetic spec", and is anchored on the reference's call
 * sites (engine_context.py:44-204) and on decoded-byte results (FIPS-197).  The CUDA engine must
 * match this file bit-for-bit on every integer primitive under identical parameters and seed.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Plain C, gcc, unsigned __int128, OpenMP over limbs.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef int64_t i64;

/* ------------------------------------------------------------------ modular arithmetic */
static inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }
static inline u64 addmod(u64 a, u64 b, u64 q) { u64 s = a + b; return s >= q ? s - q : s; }
static inline u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }

u64 ref_mulmod(u64 a, u64 b, u64 q) { return mulmod(a, b, q); }

/* Barrett reduction of a 128-bit product for the hot loops (q < 2^62): the same canonical residue as `% q` -- every result
 * of this file is unchanged -- at a few cycles instead of a 128-bit division (so that the CPU baseline timed by bench.py is
 * not a strawman).  mu = floor(2^(k+63) / q), k = bitlen(q); the quotient estimate is low by at most 3. */
typedef struct { u64 q, mu; int k1; } bar_t;
static inline bar_t bar_init(u64 q) {
    int k = 64 - __builtin_clzll(q);
    bar_t b = {q, (u64)((((u128)1) << (k + 63)) / q), k - 1};
    return b;
}
static inline u64 bar_reduce(u128 z, bar_t b) {
    u64 x = (u64)(z >> b.k1);
    u64 qh = (u64)(((u128)x * b.mu) >> 64);
    u64 r = (u64)z - qh * b.q;
    while (r >= b.q) r -= b.q;
    return r;
}
static inline u64 bmul(u64 a, u64 b, bar_t m) { return bar_reduce((u128)a * b, m); }

u64 ref_powmod(u64 a, u64 e, u64 q) {
    u64 r = 1 % q;
    a %= q;
    while (e) {
        if (e & 1) r = mulmod(r, a, q);
        a = mulmod(a, a, q);
        e >>= 1;
    }
    return r;
}

static u64 bitrev(u64 x, int bits) {
    u64 r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* spec S2: psi = x^((q-1)/2N) for the first x = 2,3,... whose power has order exactly 2N */
u64 ref_find_psi(u64 q, int logn) {
    u64 twoN = 2ull << logn, N = 1ull << logn;
    for (u64 x = 2;; x++) {
        u64 r = ref_powmod(x, (q - 1) / twoN, q);
        if (ref_powmod(r, N, q) == q - 1) return r;
    }
}

/* spec S3: tab[k] = psi^bitrev(k, logn); itab[k] = tab[k]^-1 */
void ref_ntt_tables(int logn, u64 q, u64 psi, u64 *tab, u64 *itab) {
    u64 N = 1ull << logn;
    u64 ipsi = ref_powmod(psi, q - 2, q);
    u64 p = 1, ip = 1;
    for (u64 e = 0; e < N; e++) {
        u64 k = bitrev(e, logn);
        tab[k] = p;
        itab[k] = ip;
        p = mulmod(p, psi, q);
        ip = mulmod(ip, ipsi, q);
    }
}

/* negacyclic forward NTT, natural order in, bit-reversed order out: A[k] = a(psi^(2 bitrev(k)+1)) */
void ref_ntt_fwd(u64 *a, int logn, u64 q, const u64 *tab) {
    u64 N = 1ull << logn, t = N;
    const bar_t B = bar_init(q);
    for (u64 m = 1; m < N; m <<= 1) {
        t >>= 1;
        for (u64 i = 0; i < m; i++) {
            u64 W = tab[m + i], j1 = 2 * i * t;
            for (u64 j = j1; j < j1 + t; j++) {
                u64 U = a[j], V = bmul(a[j + t], W, B);
                a[j] = addmod(U, V, q);
                a[j + t] = submod(U, V, q);
            }
        }
    }
}

void ref_ntt_inv(u64 *a, int logn, u64 q, const u64 *itab) {
    u64 N = 1ull << logn, t = 1;
    const bar_t B = bar_init(q);
    for (u64 m = N; m > 1; m >>= 1) {
        u64 h = m >> 1, j1 = 0;
        for (u64 i = 0; i < h; i++) {
            u64 W = itab[h + i];
            for (u64 j = j1; j < j1 + t; j++) {
                u64 U = a[j], V = a[j + t];
                a[j] = addmod(U, V, q);
                a[j + t] = bmul(submod(U, V, q), W, B);
            }
            j1 += 2 * t;
        }
        t <<= 1;
    }
    u64 ninv = ref_powmod(N % q, q - 2, q);
    for (u64 j = 0; j < N; j++) a[j] = bmul(a[j], ninv, B);
}

/* batches: a is [nl][N]; qs[nl]; tabs is [nl][N] */
void ref_ntt_fwd_batch(u64 *a, int nl, int logn, const u64 *qs, const u64 *tabs) {
    u64 N = 1ull << logn;
#pragma omp parallel for schedule(dynamic)
    for (int l = 0; l < nl; l++) ref_ntt_fwd(a + (size_t)l * N, logn, qs[l], tabs + (size_t)l * N);
}
void ref_ntt_inv_batch(u64 *a, int nl, int logn, const u64 *qs, const u64 *itabs) {
    u64 N = 1ull << logn;
#pragma omp parallel for schedule(dynamic)
    for (int l = 0; l < nl; l++) ref_ntt_inv(a + (size_t)l * N, logn, qs[l], itabs + (size_t)l * N);
}

/* ------------------------------------------------------------------ limb-wise element ops */
/* out[l][k] = a[l][k] * b[l][k]  (+ acc) */
void ref_mul_batch(u64 *out, const u64 *a, const u64 *b, int nl, size_t N, const u64 *qs) {
#pragma omp parallel for schedule(static)
    for (int l = 0; l < nl; l++) {
        const bar_t B = bar_init(qs[l]);
        for (size_t k = 0; k < N; k++) out[l * N + k] = bmul(a[l * N + k], b[l * N + k], B);
    }
}
void ref_muladd_batch(u64 *acc, const u64 *a, const u64 *b, int nl, size_t N, const u64 *qs) {
#pragma omp parallel for schedule(static)
    for (int l = 0; l < nl; l++) {
        const bar_t B = bar_init(qs[l]);
        for (size_t k = 0; k < N; k++)
            acc[l * N + k] = addmod(acc[l * N + k], bmul(a[l * N + k], b[l * N + k], B), qs[l]);
    }
}
void ref_add_batch(u64 *out, const u64 *a, const u64 *b, int nl, size_t N, const u64 *qs) {
#pragma omp parallel for schedule(static)
    for (int l = 0; l < nl; l++)
        for (size_t k = 0; k < N; k++) out[l * N + k] = addmod(a[l * N + k], b[l * N + k], qs[l]);
}
void ref_sub_batch(u64 *out, const u64 *a, const u64 *b, int nl, size_t N, const u64 *qs) {
#pragma omp parallel for schedule(static)
    for (int l = 0; l < nl; l++)
        for (size_t k = 0; k < N; k++) out[l * N + k] = submod(a[l * N + k], b[l * N + k], qs[l]);
}
/* out[l][k] = a[l][k] * s[l] */
void ref_mul_scalar_batch(u64 *out, const u64 *a, const u64 *s, int nl, size_t N, const u64 *qs) {
#pragma omp parallel for schedule(static)
    for (int l = 0; l < nl; l++) {
        const bar_t B = bar_init(qs[l]);
        for (size_t k = 0; k < N; k++) out[l * N + k] = bmul(a[l * N + k], s[l], B);
    }
}
/* NTT-domain multiply by the 2-term polynomial R + I*X^(N/2): first half of the bit-reversed array sees
 * cp[l] = R + I*J, second half cm[l] = R - I*J  (J = psi^(N/2)); DESIGN.md spec S7 */
void ref_mul_const_batch(u64 *out, const u64 *a, const u64 *cp, const u64 *cm, int nl, size_t N, const u64 *qs) {
#pragma omp parallel for schedule(static)
    for (int l = 0; l < nl; l++) {
        const bar_t B = bar_init(qs[l]);
        for (size_t k = 0; k < N; k++)
            out[l * N + k] = bmul(a[l * N + k], k < N / 2 ? cp[l] : cm[l], B);
    }
}
/* gather: out[l][k] = a[l][perm[k]] */
void ref_permute_batch(u64 *out, const u64 *a, const uint32_t *perm, int nl, size_t N) {
#pragma omp parallel for schedule(static)
    for (int l = 0; l < nl; l++)
        for (size_t k = 0; k < N; k++) out[l * N + k] = a[l * N + perm[k]];
}

/* spec S5 (fast basis conversion, no overflow correction):
 * out[t][k] = sum_i ((in[i][k] * hatinv[i]) mod sq[i]) * hat[i][t]   mod tq[t] */
void ref_baseconv(u64 *out, const u64 *in, size_t N, int ns, const u64 *sq, const u64 *hatinv, int nt,
                  const u64 *tq, const u64 *hat /* [ns][nt] */) {
    bar_t BS[64], BT[64];
    for (int i = 0; i < ns; i++) BS[i] = bar_init(sq[i]);
    for (int t = 0; t < nt; t++) BT[t] = bar_init(tq[t]);
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < N; k++) {
        u64 y[64];
        for (int i = 0; i < ns; i++) y[i] = bmul(in[i * N + k], hatinv[i], BS[i]);
        for (int t = 0; t < nt; t++) {
            u64 acc = 0, q = tq[t];
            /* (y mod q) * hat mod q == y * hat mod q; y * hat < 2^124 fits the Barrett input range */
            for (int i = 0; i < ns; i++) acc = addmod(acc, bmul(y[i] % q, hat[i * nt + t], BT[t]), q);
            out[t * N + k] = acc;
        }
    }
}

/* spec S5' (exact, centred conversion used by ModDown): as ref_baseconv, then the overflow count
 * u = round(sum_i y_i / s_i) (fp64, sequential, separately rounded products) times D = prod s_i is taken out:
 * out[t] = sum_i y_i hat[i][t] - u D  (mod tq[t]);  negD[t] = tq[t] - (D mod tq[t]) */
void ref_baseconv_exact(u64 *out, const u64 *in, size_t N, int ns, const u64 *sq, const u64 *hatinv, int nt,
                        const u64 *tq, const u64 *hat /* [ns][nt] */, const double *inv_src, const u64 *negD) {
    bar_t BS[64], BT[64];
    for (int i = 0; i < ns; i++) BS[i] = bar_init(sq[i]);
    for (int t = 0; t < nt; t++) BT[t] = bar_init(tq[t]);
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < N; k++) {
        u64 y[64];
        double v = 0.0;
        for (int i = 0; i < ns; i++) {
            y[i] = bmul(in[i * N + k], hatinv[i], BS[i]);
            volatile double p = (double)y[i] * inv_src[i];
            v = v + p;
        }
        u64 u = (u64)(i64)nearbyint(v);
        for (int t = 0; t < nt; t++) {
            u64 acc = 0, q = tq[t];
            for (int i = 0; i < ns; i++) acc = addmod(acc, bmul(y[i] % q, hat[i * nt + t], BT[t]), q);
            acc = addmod(acc, bmul(u % q, negD[t], BT[t]), q);
            out[t * N + k] = acc;
        }
    }
}

/* signed small coefficients -> residues, [nl][N] */
void ref_reduce_i64_batch(u64 *out, const i64 *v, int nl, size_t N, const u64 *qs) {
#pragma omp parallel for schedule(static)
    for (int l = 0; l < nl; l++) {
        u64 q = qs[l];
        for (size_t k = 0; k < N; k++) {
            i64 x = v[k];
            u64 r = (u64)(x < 0 ? -x : x) % q;
            out[l * N + k] = (x < 0 && r) ? q - r : r;
        }
    }
}

/* ------------------------------------------------------------------ sampling (spec S8) */
static inline u64 mix64(u64 z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
u64 ref_rand64(u64 seed, u64 stream, u64 idx) { return mix64(mix64(seed + stream * 0xD1342543DE82EF95ull) + idx); }

/* uniform in [0,q): 8-try rejection on counter (limb*N + k)*8 + t */
void ref_sample_uniform(u64 *out, size_t N, u64 q, u64 seed, u64 stream, u64 limb_index) {
    u64 bound = (u64)0 - ((u64)0 - q) % q; /* largest multiple of q that fits in 2^64, mod 2^64 (0 means 2^64) */
    for (size_t k = 0; k < N; k++) {
        u64 r = 0;
        for (int t = 0; t < 8; t++) {
            r = ref_rand64(seed, stream, (limb_index * N + k) * 8 + t);
            if (bound == 0 || r < bound) break;
        }
        out[k] = r % q;
    }
}
/* centred binomial, 21+21 bits: variance 10.5 (sigma 3.24) */
void ref_sample_cbd(i64 *out, size_t N, u64 seed, u64 stream) {
    for (size_t k = 0; k < N; k++) {
        u64 r = ref_rand64(seed, stream, k);
        out[k] = (i64)__builtin_popcountll(r & 0x1FFFFF) - (i64)__builtin_popcountll((r >> 21) & 0x1FFFFF);
    }
}
/* {-1,0,1} with probabilities 1/4,1/2,1/4 */
void ref_sample_ternary(i64 *out, size_t N, u64 seed, u64 stream) {
    for (size_t k = 0; k < N; k++) {
        u64 r = ref_rand64(seed, stream, k);
        out[k] = (i64)(r & 1) - (i64)((r >> 1) & 1);
    }
}
/* sparse ternary secret, exactly h non-zeros: partial Fisher-Yates on stream `stream` */
void ref_sample_sparse(i64 *out, size_t N, int h, u64 seed, u64 stream) {
    uint32_t *perm = (uint32_t *)malloc(N * sizeof(uint32_t));
    for (size_t k = 0; k < N; k++) { perm[k] = (uint32_t)k; out[k] = 0; }
    for (int i = 0; i < h; i++) {
        u64 j = i + ref_rand64(seed, stream, (u64)i) % (N - i);
        uint32_t tmp = perm[i]; perm[i] = perm[j]; perm[j] = tmp;
        out[perm[i]] = (ref_rand64(seed, stream, (u64)h + i) & 1) ? 1 : -1;
    }
    free(perm);
}

/* ------------------------------------------------------------------ canonical embedding (spec S9) */
/* tables: rot[j] = 5^j mod 2N (j < N/2), ksi[k] = exp(2 pi i k / 2N) as (re,im) pairs, k <= 2N */
void ref_fft_tables(int logn, uint32_t *rot, double *ksi) {
    u64 N = 1ull << logn, M = 2 * N, n = N / 2;
    u64 p = 1;
    for (u64 j = 0; j < n; j++) { rot[j] = (uint32_t)p; p = (p * 5) % M; }
    for (u64 k = 0; k <= M; k++) {
        double ang = 2.0 * M_PI * (double)k / (double)M;
        ksi[2 * k] = cos(ang);
        ksi[2 * k + 1] = sin(ang);
    }
}
static void bitrev_permute_c(double *v, u64 n, int bits) {
    for (u64 i = 0; i < n; i++) {
        u64 j = bitrev(i, bits);
        if (i < j) {
            double a = v[2 * i], b = v[2 * i + 1];
            v[2 * i] = v[2 * j]; v[2 * i + 1] = v[2 * j + 1];
            v[2 * j] = a; v[2 * j + 1] = b;
        }
    }
}
/* decode direction: w (coefficient pairs) -> z (slots); v is n complex (re,im) doubles, in place.
 * complex product (a+bi)(c+di) = (ac - bd) + (ad + bc) i with separately rounded products (no FMA). */
void ref_special_fft(double *v, int logn, const uint32_t *rot, const double *ksi) {
    u64 N = 1ull << logn, M = 2 * N, n = N / 2;
    bitrev_permute_c(v, n, logn - 1);
    for (u64 len = 2; len <= n; len <<= 1) {
        u64 lenh = len >> 1, lenq = len << 2, gap = M / lenq;
        for (u64 i = 0; i < n; i += len)
            for (u64 j = 0; j < lenh; j++) {
                u64 idx = (rot[j] % lenq) * gap;
                double wr = ksi[2 * idx], wi = ksi[2 * idx + 1];
                double ur = v[2 * (i + j)], ui = v[2 * (i + j) + 1];
                double xr = v[2 * (i + j + lenh)], xi = v[2 * (i + j + lenh) + 1];
                volatile double p0 = xr * wr, p1 = xi * wi, p2 = xr * wi, p3 = xi * wr;
                double tr = p0 - p1, ti = p2 + p3;
                v[2 * (i + j)] = ur + tr; v[2 * (i + j) + 1] = ui + ti;
                v[2 * (i + j + lenh)] = ur - tr; v[2 * (i + j + lenh) + 1] = ui - ti;
            }
    }
}
/* encode direction: z -> w, includes the 1/n scaling */
void ref_special_ifft(double *v, int logn, const uint32_t *rot, const double *ksi) {
    u64 N = 1ull << logn, M = 2 * N, n = N / 2;
    for (u64 len = n; len >= 2; len >>= 1) {
        u64 lenh = len >> 1, lenq = len << 2, gap = M / lenq;
        for (u64 i = 0; i < n; i += len)
            for (u64 j = 0; j < lenh; j++) {
                u64 idx = (lenq - (rot[j] % lenq)) * gap;
                double wr = ksi[2 * idx], wi = ksi[2 * idx + 1];
                double ar = v[2 * (i + j)], ai = v[2 * (i + j) + 1];
                double br = v[2 * (i + j + lenh)], bi = v[2 * (i + j + lenh) + 1];
                double dr = ar - br, di = ai - bi;
                volatile double p0 = dr * wr, p1 = di * wi, p2 = dr * wi, p3 = di * wr;
                v[2 * (i + j)] = ar + br; v[2 * (i + j) + 1] = ai + bi;
                v[2 * (i + j + lenh)] = p0 - p1; v[2 * (i + j + lenh) + 1] = p2 + p3;
            }
    }
    bitrev_permute_c(v, n, logn - 1);
    double inv = 1.0 / (double)n;
    for (u64 k = 0; k < 2 * n; k++) v[k] *= inv;
}
/* w (n complex) scaled by `scale`, rounded half-to-even -> N signed coefficients (m_k = Re w_k, m_{k+n} = Im w_k) */
int ref_round_coeffs(i64 *out, const double *w, size_t n, double scale) {
    int overflow = 0;
    for (size_t k = 0; k < n; k++) {
        volatile double a = w[2 * k] * scale, b = w[2 * k + 1] * scale;
        double ra = nearbyint(a), rb = nearbyint(b);
        if (fabs(ra) >= 4.0e18 || fabs(rb) >= 4.0e18) overflow = 1;
        out[k] = (i64)ra;
        out[k + n] = (i64)rb;
    }
    return overflow;
}
/* centred lift of one limb -> doubles divided by scale, packed as w */
void ref_center_to_w(double *w, const u64 *coef, size_t n, u64 q, double scale) {
    u64 half = q >> 1;
    for (size_t k = 0; k < n; k++) {
        u64 a = coef[k], b = coef[k + n];
        double da = a > half ? -(double)(q - a) : (double)a;
        double db = b > half ? -(double)(q - b) : (double)b;
        w[2 * k] = da / scale;
        w[2 * k + 1] = db / scale;
    }
}
