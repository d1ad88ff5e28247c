// ntt.cu -- negacyclic NTT / iNTT, 64-bit RNS limbs, sm_100a.  N = 2^16 (product) and 2^12 (tests).
//
// Layout of the transform.  The length-N Cooley-Tukey NTT with the psi-powers table in
// bit-reversed order (DESIGN.md spec S3, same table as oracle/ckks_ref.c:ref_ntt_tables) is
// split into two passes over the limb viewed as an R x 256 row-major matrix (R = N/256):
//
//   pass A  stages 1..log R          butterflies couple whole rows: "columns"
//   pass B  stages log R+1..log N    butterflies stay inside one contiguous 256-element row
//
// No transposition and no extra inter-pass twiddle multiply is needed in this formulation:
// the bit-reversed table is self-similar, a radix-16 sub-transform rooted at table index X
// uses entries (X << s) + g, s = 0..3 (forward) -- for N = 2^16: X = 1, 16+rh, 256+R, 4096+16R+jh
// for the four register rounds.  Each thread keeps 16 coefficients in registers for 4 stages,
// with ONE shared-memory exchange per pass.  Global accesses are 128-byte coalesced in both
// passes; pass B stages its stores through shared memory (padded 1-in-16 so the 64-bit accesses
// are bank-conflict free).
//
// Two arithmetic paths, chosen per limb (a CTA works on one limb, so the choice is uniform):
//   * integer path (the 60/61-bit moduli q_0 and the special primes): Harvey lazy butterflies with Shoup
//     twiddles, values in [0,4q) forward / [0,2q) inverse.  One Shoup multiplication costs a 64x64 high
//     product: measured 3.86 modmul/clk/SM on the B200 (profiles/r1_pipe_peaks.txt).
//   * FP64 path (the ~2^50 scale primes, q < 1.4 * 2^50): coefficients are held as exact integers in doubles and
//     multiplied with two DMUL, two DFMA, one rounding and one add (error-free product + quotient estimate):
//     measured 7.70 modmul/clk/SM, on the FP64 pipe, which the integer path leaves idle.  Values stay two-sided
//     lazy (|x| < 2^53) and are folded to |x| <= q/2 every two stages.  Every operation is exact, so the
//     canonical output is bit-identical to the integer path and to the oracle.
//
// Algorithmic bytes: 2*N*8 = 1 MiB per limb-NTT at N = 2^16; measured DRAM traffic is 1.63x that
// (the twiddle tables are as large as the data).
#include "ntt.cuh"

namespace {

constexpr int kThreads = 256;
#ifndef NTT_MIN_BLOCKS
#define NTT_MIN_BLOCKS 3      // register cap 80, no spills; measured best of {2,3,4,5} (profiles/r1_ntt_occupancy_sweep.txt)
#endif
#ifndef NTT_FP64
#define NTT_FP64 1            // 0: every limb takes the integer path (A/B measurement)
#endif
// FP64 path bound: |x| < 4.6 q must stay below 2^53
#define NTT_FP_LIMIT 1576258512130867ull     /* 1.4 * 2^50 */

__device__ __forceinline__ int pad16(int i) { return i + (i >> 4); }
__device__ __forceinline__ bool use_fp(u64 q) { return NTT_FP64 && q < NTT_FP_LIMIT; }

// ============================================================================================ integer path
// Forward radix-16 block rooted at table index X: 4 CT stages on x[0..15], lazy in [0,4q).
__device__ __forceinline__ void fwd16(u64 (&x)[16], u32 X, const u64* __restrict__ W, const u64* __restrict__ Ws,
                                      u64 q) {
    const u64 q2 = 2 * q;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int span = 8 >> s;
        u64 w[8], ws[8];
#pragma unroll
        for (int g = 0; g < 8; g++)
            if (g < (1 << s)) { w[g] = ldg(W + (X << s) + g); ws[g] = ldg(Ws + (X << s) + g); }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int g = i / span, k0 = g * 2 * span + (i % span), k1 = k0 + span;
            u64 u = x[k0];
            u = u >= q2 ? u - q2 : u;
            const u64 v = shoup_mul_lazy(x[k1], w[g], ws[g], q);
            x[k0] = u + v;
            x[k1] = u - v + q2;
        }
    }
}

// Inverse radix-16 block rooted at X: 4 GS stages, values kept in [0,2q).
// If FINAL, the last stage folds in N^-1 (scaling the sum by ninv and the twiddle by ninv).
template <bool FINAL>
__device__ __forceinline__ void inv16(u64 (&x)[16], u32 X, const u64* __restrict__ W, const u64* __restrict__ Ws,
                                      u64 q, u64 ninv, u64 ninv_s, u64 w1n, u64 w1n_s) {
    const u64 q2 = 2 * q;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int span = 1 << s;
        u64 w[8], ws[8];
#pragma unroll
        for (int g = 0; g < 8; g++)
            if (g < (8 >> s)) {
                if (FINAL && s == 3) { w[g] = w1n; ws[g] = w1n_s; }
                else { w[g] = ldg(W + (X << (3 - s)) + g); ws[g] = ldg(Ws + (X << (3 - s)) + g); }
            }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int g = i / span, k0 = g * 2 * span + (i % span), k1 = k0 + span;
            const u64 u = x[k0], v = x[k1];
            u64 sum = u + v;
            sum = sum >= q2 ? sum - q2 : sum;
            if (FINAL && s == 3) sum = shoup_mul_lazy(sum, ninv, ninv_s, q);
            x[k0] = sum;
            x[k1] = shoup_mul_lazy(u - v + q2, w[g], ws[g], q);
        }
    }
}

__device__ __forceinline__ u64 canon4(u64 v, u64 q) {
    v = v >= 2 * q ? v - 2 * q : v;
    return v >= q ? v - q : v;
}
__device__ __forceinline__ u64 canon2(u64 v, u64 q) { return v >= q ? v - q : v; }

// ============================================================================================ FP64 path
// a * w mod q as an exact integer in (-2q, 2q): a any integer with |a| < 2^53, w < q < 2^51, wq = fl(w / q).
//   h = fl(a w), l = a w - h (exact, FMA), c = rint(a wq) (quotient, off by at most 2), r = (h - c q) + l (both exact)
__device__ __forceinline__ double modmul_fp(double a, double w, double wq, double q) {
    const double h = fmul_rn(a, w);
    const double l = ffma_rn(a, w, -h);
    const double c = frint(fmul_rn(a, wq));
    return fadd_rn(ffma_rn(-c, q, h), l);
}
// fold a lazy value (|x| < 2^53) to |x| <= q/2 (+ one q when the quotient estimate is off by one)
__device__ __forceinline__ double fold_fp(double x, double q, double qinv) {
    return ffma_rn(-frint(fmul_rn(x, qinv)), q, x);
}
// exact canonical residue in [0,q) as an integer
__device__ __forceinline__ u64 canon_fp(double x, double q, double qinv) {
    x = fold_fp(x, q, qinv);
    x = x < 0.0 ? fadd_rn(x, q) : x;
    x = x >= q ? fsub_rn(x, q) : x;
    return (u64)d2ll_rn(x);
}

struct FpMod {
    double q, qinv, ninv, ninvq, w1n, w1nq;
};

// Forward radix-16 block, 4 CT stages.  In: |x| <= 0.51 q.  Out: |x| < 4.6 q after two stages, folded, then again:
// the caller receives |x| < 4.6 q (lazy) -- two folds per block keep everything below 2^53.
__device__ __forceinline__ void fwd16_fp(double (&x)[16], u32 X, const double* __restrict__ W,
                                         const double* __restrict__ Wq, const FpMod& m) {
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int span = 8 >> s;
        if (s == 2) {
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = fold_fp(x[k], m.q, m.qinv);
        }
        double w[8], wq[8];
#pragma unroll
        for (int g = 0; g < 8; g++)
            if (g < (1 << s)) { w[g] = ldg(W + (X << s) + g); wq[g] = ldg(Wq + (X << s) + g); }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int g = i / span, k0 = g * 2 * span + (i % span), k1 = k0 + span;
            const double u = x[k0];
            const double v = modmul_fp(x[k1], w[g], wq[g], m.q);
            x[k0] = fadd_rn(u, v);
            x[k1] = fsub_rn(u, v);
        }
    }
}
// Inverse radix-16 block, 4 GS stages.  In: |x| <= 0.51 q; sums double per stage, so fold after two stages.
template <bool FINAL>
__device__ __forceinline__ void inv16_fp(double (&x)[16], u32 X, const double* __restrict__ W,
                                         const double* __restrict__ Wq, const FpMod& m) {
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int span = 1 << s;
        if (s == 2) {
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = fold_fp(x[k], m.q, m.qinv);
        }
        double w[8], wq[8];
#pragma unroll
        for (int g = 0; g < 8; g++)
            if (g < (8 >> s)) {
                if (FINAL && s == 3) { w[g] = m.w1n; wq[g] = m.w1nq; }
                else { w[g] = ldg(W + (X << (3 - s)) + g); wq[g] = ldg(Wq + (X << (3 - s)) + g); }
            }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int g = i / span, k0 = g * 2 * span + (i % span), k1 = k0 + span;
            const double u = x[k0], v = x[k1];
            double sum = fadd_rn(u, v);
            if (FINAL && s == 3) sum = modmul_fp(sum, m.ninv, m.ninvq, m.q);
            x[k0] = sum;
            x[k1] = modmul_fp(fsub_rn(u, v), w[g], wq[g], m.q);
        }
    }
}

__device__ __forceinline__ FpMod fp_mod(const ModConst& mc) {
    FpMod m;
    m.q = ull2d_rn(mc.q);
    m.qinv = fdiv_rn(1.0, m.q);
    m.ninv = ull2d_rn(mc.ninv);
    m.ninvq = fdiv_rn(m.ninv, m.q);
    m.w1n = ull2d_rn(mc.w1n);
    m.w1nq = fdiv_rn(m.w1n, m.q);
    return m;
}
// lazy doubles travel between the two passes in the u64 buffer as raw bits
__device__ __forceinline__ u64 d2bits(double x) { union { double d; u64 u; } c; c.d = x; return c.u; }
__device__ __forceinline__ double bits2d(u64 x) { union { double d; u64 u; } c; c.u = x; return c.d; }

// ============================================================================================ forward, pass A
// LOGR = 8: R = 256 rows, tile = 16 columns x 256 rows, two radix-16 rounds (X = 1, then 16 + rr).
// LOGR = 4: R = 16 rows,  tile = 256 columns x 16 rows, one radix-16 round (X = 1).
template <int LOGR, bool FP>
__device__ __forceinline__ void fwd_passA_body(const u64* __restrict__ src, u64* __restrict__ dst, u64* sm, int limb,
                                               int slimb, int tile, size_t N, const ModConst& mc, const NttTables& T,
                                               int mod) {
    constexpr int RG = (1 << LOGR) / 16;          // row groups per column: 16 or 1
    constexpr int TC = kThreads / RG;             // columns per CTA: 16 or 256
    const u64 q = mc.q;
    const u64* W = T.fwd + (size_t)mod * N;
    const u64* Ws = T.fwd_s + (size_t)mod * N;
    const double* Wd = T.fwd_d + (size_t)mod * N;
    const double* Wq = T.fwd_q + (size_t)mod * N;
    const FpMod fm = fp_mod(mc);
    double* smd = reinterpret_cast<double*>(sm);
    if (LOGR == 8) {
        FOR_THREADS {
            const int c = threadIdx.x % TC, rr = threadIdx.x / TC;
            const size_t sbase = (size_t)slimb * N + tile * TC + c;
            if (FP) {
                double x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = ull2d_rn(src[sbase + (size_t)(rr + 16 * k) * 256]);
                fwd16_fp(x, 1u, Wd, Wq, fm);
#pragma unroll
                for (int k = 0; k < 16; k++) smd[(rr + 16 * k) * TC + c] = fold_fp(x[k], fm.q, fm.qinv);
            } else {
                u64 x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = src[sbase + (size_t)(rr + 16 * k) * 256];
                fwd16(x, 1u, W, Ws, q);
#pragma unroll
                for (int k = 0; k < 16; k++) sm[(rr + 16 * k) * TC + c] = x[k];
            }
        }
        BLOCK_SYNC;
        FOR_THREADS {
            const int c = threadIdx.x % TC, rr = threadIdx.x / TC;
            const size_t base = (size_t)limb * N + tile * TC + c;
            if (FP) {
                double x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = smd[(16 * rr + k) * TC + c];
                fwd16_fp(x, 16u + rr, Wd, Wq, fm);
#pragma unroll
                for (int k = 0; k < 16; k++) dst[base + (size_t)(16 * rr + k) * 256] = d2bits(fold_fp(x[k], fm.q, fm.qinv));
            } else {
                u64 x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = sm[(16 * rr + k) * TC + c];
                fwd16(x, 16u + rr, W, Ws, q);
#pragma unroll
                for (int k = 0; k < 16; k++) dst[base + (size_t)(16 * rr + k) * 256] = x[k];   // lazy [0,4q)
            }
        }
    } else {
        FOR_THREADS {
            const size_t base = (size_t)limb * N + threadIdx.x, sbase = (size_t)slimb * N + threadIdx.x;
            if (FP) {
                double x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = ull2d_rn(src[sbase + (size_t)k * 256]);
                fwd16_fp(x, 1u, Wd, Wq, fm);
#pragma unroll
                for (int k = 0; k < 16; k++) dst[base + (size_t)k * 256] = d2bits(fold_fp(x[k], fm.q, fm.qinv));
            } else {
                u64 x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = src[sbase + (size_t)k * 256];
                fwd16(x, 1u, W, Ws, q);
#pragma unroll
                for (int k = 0; k < 16; k++) dst[base + (size_t)k * 256] = x[k];
            }
        }
    }
}
template <int LOGR>
__global__ void __launch_bounds__(kThreads, NTT_MIN_BLOCKS)
ntt_fwd_passA(const u64* __restrict__ src, u64* __restrict__ dst, const GRID_CONST NttJob J, NttTables T) {
    CKKS_SHARED u64 sm[LOGR == 8 ? 256 * 16 : 1];
    if (J.cnt[blockIdx.z] && blockIdx.y >= J.cnt[blockIdx.z]) return;
    const int limb = J.rows[blockIdx.z][blockIdx.y], slimb = J.srows[blockIdx.z][blockIdx.y], tile = blockIdx.x;
    const int mod = J.mods[blockIdx.z][blockIdx.y];
    const size_t N = (size_t)1 << T.logn;
    src += blockIdx.z * J.szs;
    dst += blockIdx.z * J.dzs;
    const ModConst mc = T.mc[mod];
    if (use_fp(mc.q)) fwd_passA_body<LOGR, true>(src, dst, sm, limb, slimb, tile, N, mc, T, mod);
    else fwd_passA_body<LOGR, false>(src, dst, sm, limb, slimb, tile, N, mc, T, mod);
}

// ============================================================================================ forward, pass B
// 16 rows of 256 per CTA; row with global index Rg is rooted at table index R + Rg.
template <bool FP>
__device__ __forceinline__ void fwd_passB_body(u64* __restrict__ g, u64* sm, int tile, u32 Rn, const ModConst& mc,
                                               const NttTables& T, int mod, size_t N) {
    const u64 q = mc.q;
    const u64* W = T.fwd + (size_t)mod * N;
    const u64* Ws = T.fwd_s + (size_t)mod * N;
    const double* Wd = T.fwd_d + (size_t)mod * N;
    const double* Wq = T.fwd_q + (size_t)mod * N;
    const FpMod fm = fp_mod(mc);
    double* smd = reinterpret_cast<double*>(sm);
    FOR_THREADS {
        const int jj = threadIdx.x % 16, row = threadIdx.x / 16;
        const u32 R = tile * 16 + row;
        if (FP) {
            double x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = bits2d(g[row * 256 + jj + 16 * k]);
            fwd16_fp(x, Rn + R, Wd, Wq, fm);
#pragma unroll
            for (int k = 0; k < 16; k++) smd[pad16(row * 256 + jj + 16 * k)] = fold_fp(x[k], fm.q, fm.qinv);
        } else {
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = g[row * 256 + jj + 16 * k];
            fwd16(x, Rn + R, W, Ws, q);
#pragma unroll
            for (int k = 0; k < 16; k++) sm[pad16(row * 256 + jj + 16 * k)] = x[k];
        }
    }
    BLOCK_SYNC;
    FOR_THREADS {
        const int jj = threadIdx.x % 16, row = threadIdx.x / 16;
        const u32 R = tile * 16 + row;
        // each thread rewrites exactly the 16 slots it just read, so no barrier is needed before this store
        if (FP) {
            double x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = smd[pad16(row * 256 + 16 * jj + k)];
            fwd16_fp(x, 16u * (Rn + R) + jj, Wd, Wq, fm);
#pragma unroll
            for (int k = 0; k < 16; k++) sm[pad16(row * 256 + 16 * jj + k)] = canon_fp(x[k], fm.q, fm.qinv);
        } else {
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = sm[pad16(row * 256 + 16 * jj + k)];
            fwd16(x, 16u * (Rn + R) + jj, W, Ws, q);
#pragma unroll
            for (int k = 0; k < 16; k++) sm[pad16(row * 256 + 16 * jj + k)] = canon4(x[k], q);
        }
    }
    BLOCK_SYNC;
    FOR_THREADS {
#pragma unroll
        for (int k = 0; k < 16; k++) g[k * 256 + threadIdx.x] = sm[pad16(k * 256 + threadIdx.x)];
    }
}
__global__ void __launch_bounds__(kThreads, NTT_MIN_BLOCKS)
ntt_fwd_passB(u64* __restrict__ data, const GRID_CONST NttJob J, NttTables T) {
    CKKS_SHARED u64 sm[16 * 256 + 16 * 16];
    if (J.cnt[blockIdx.z] && blockIdx.y >= J.cnt[blockIdx.z]) return;
    const int limb = J.rows[blockIdx.z][blockIdx.y], tile = blockIdx.x;
    const int mod = J.mods[blockIdx.z][blockIdx.y];
    const size_t N = (size_t)1 << T.logn;
    const u32 Rn = (u32)(N >> 8);
    data += blockIdx.z * J.dzs;
    const ModConst mc = T.mc[mod];
    u64* g = data + (size_t)limb * N + (size_t)tile * 16 * 256;
    if (use_fp(mc.q)) fwd_passB_body<true>(g, sm, tile, Rn, mc, T, mod, N);
    else fwd_passB_body<false>(g, sm, tile, Rn, mc, T, mod, N);
}

// ============================================================================================ inverse, pass B^-1
template <bool FP>
__device__ __forceinline__ void inv_passB_body(const u64* __restrict__ s_in, u64* __restrict__ d_out, u64* sm, int tile,
                                               u32 Rn, const ModConst& mc, const NttTables& T, int mod, size_t N) {
    const u64 q = mc.q;
    const u64* W = T.inv + (size_t)mod * N;
    const u64* Ws = T.inv_s + (size_t)mod * N;
    const double* Wd = T.inv_d + (size_t)mod * N;
    const double* Wq = T.inv_q + (size_t)mod * N;
    const FpMod fm = fp_mod(mc);
    double* smd = reinterpret_cast<double*>(sm);
    FOR_THREADS {
#pragma unroll
        for (int k = 0; k < 16; k++) sm[pad16(k * 256 + threadIdx.x)] = s_in[k * 256 + threadIdx.x];
    }
    BLOCK_SYNC;
    FOR_THREADS {
        const int jj = threadIdx.x % 16, row = threadIdx.x / 16;
        const u32 R = tile * 16 + row;
        if (FP) {
            double x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = ull2d_rn(sm[pad16(row * 256 + 16 * jj + k)]);
            inv16_fp<false>(x, 16u * (Rn + R) + jj, Wd, Wq, fm);
#pragma unroll
            for (int k = 0; k < 16; k++) smd[pad16(row * 256 + 16 * jj + k)] = fold_fp(x[k], fm.q, fm.qinv);
        } else {
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = sm[pad16(row * 256 + 16 * jj + k)];
            inv16<false>(x, 16u * (Rn + R) + jj, W, Ws, q, 0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < 16; k++) sm[pad16(row * 256 + 16 * jj + k)] = x[k];
        }
    }
    BLOCK_SYNC;
    FOR_THREADS {
        const int jj = threadIdx.x % 16, row = threadIdx.x / 16;
        const u32 R = tile * 16 + row;
        if (FP) {
            double x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = smd[pad16(row * 256 + jj + 16 * k)];
            inv16_fp<false>(x, Rn + R, Wd, Wq, fm);
#pragma unroll
            for (int k = 0; k < 16; k++) d_out[row * 256 + jj + 16 * k] = d2bits(fold_fp(x[k], fm.q, fm.qinv));
        } else {
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = sm[pad16(row * 256 + jj + 16 * k)];
            inv16<false>(x, Rn + R, W, Ws, q, 0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < 16; k++) d_out[row * 256 + jj + 16 * k] = x[k];       // lazy [0,2q)
        }
    }
}
__global__ void __launch_bounds__(kThreads, NTT_MIN_BLOCKS)
ntt_inv_passB(const u64* __restrict__ src, u64* __restrict__ dst, const GRID_CONST NttJob J, NttTables T) {
    CKKS_SHARED u64 sm[16 * 256 + 16 * 16];
    if (J.cnt[blockIdx.z] && blockIdx.y >= J.cnt[blockIdx.z]) return;
    const int limb = J.rows[blockIdx.z][blockIdx.y], slimb = J.srows[blockIdx.z][blockIdx.y], tile = blockIdx.x;
    const int mod = J.mods[blockIdx.z][blockIdx.y];
    const size_t N = (size_t)1 << T.logn;
    const u32 Rn = (u32)(N >> 8);
    src += blockIdx.z * J.szs;
    dst += blockIdx.z * J.dzs;
    const ModConst mc = T.mc[mod];
    const u64* s_in = src + (size_t)slimb * N + (size_t)tile * 16 * 256;
    u64* d_out = dst + (size_t)limb * N + (size_t)tile * 16 * 256;
    if (use_fp(mc.q)) inv_passB_body<true>(s_in, d_out, sm, tile, Rn, mc, T, mod, N);
    else inv_passB_body<false>(s_in, d_out, sm, tile, Rn, mc, T, mod, N);
}

// ============================================================================================ inverse, pass A^-1
template <int LOGR, bool FP>
__device__ __forceinline__ void inv_passA_body(u64* __restrict__ data, u64* sm, int limb, int tile, size_t N,
                                               const ModConst& mc, const NttTables& T, int mod) {
    constexpr int RG = (1 << LOGR) / 16;
    constexpr int TC = kThreads / RG;
    const u64 q = mc.q;
    const u64* W = T.inv + (size_t)mod * N;
    const u64* Ws = T.inv_s + (size_t)mod * N;
    const double* Wd = T.inv_d + (size_t)mod * N;
    const double* Wq = T.inv_q + (size_t)mod * N;
    const FpMod fm = fp_mod(mc);
    double* smd = reinterpret_cast<double*>(sm);
    if (LOGR == 8) {
        FOR_THREADS {
            const int c = threadIdx.x % TC, rr = threadIdx.x / TC;
            const size_t base = (size_t)limb * N + tile * TC + c;
            if (FP) {
                double x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = bits2d(data[base + (size_t)(16 * rr + k) * 256]);
                inv16_fp<false>(x, 16u + rr, Wd, Wq, fm);
#pragma unroll
                for (int k = 0; k < 16; k++) smd[(16 * rr + k) * TC + c] = fold_fp(x[k], fm.q, fm.qinv);
            } else {
                u64 x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = data[base + (size_t)(16 * rr + k) * 256];
                inv16<false>(x, 16u + rr, W, Ws, q, 0, 0, 0, 0);
#pragma unroll
                for (int k = 0; k < 16; k++) sm[(16 * rr + k) * TC + c] = x[k];
            }
        }
        BLOCK_SYNC;
        FOR_THREADS {
            const int c = threadIdx.x % TC, rr = threadIdx.x / TC;
            const size_t base = (size_t)limb * N + tile * TC + c;
            if (FP) {
                double x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = smd[(rr + 16 * k) * TC + c];
                inv16_fp<true>(x, 1u, Wd, Wq, fm);
#pragma unroll
                for (int k = 0; k < 16; k++) data[base + (size_t)(rr + 16 * k) * 256] = canon_fp(x[k], fm.q, fm.qinv);
            } else {
                u64 x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = sm[(rr + 16 * k) * TC + c];
                inv16<true>(x, 1u, W, Ws, q, mc.ninv, mc.ninv_s, mc.w1n, mc.w1n_s);
#pragma unroll
                for (int k = 0; k < 16; k++) data[base + (size_t)(rr + 16 * k) * 256] = canon2(x[k], q);
            }
        }
    } else {
        FOR_THREADS {
            const size_t base = (size_t)limb * N + threadIdx.x;
            if (FP) {
                double x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = bits2d(data[base + (size_t)k * 256]);
                inv16_fp<true>(x, 1u, Wd, Wq, fm);
#pragma unroll
                for (int k = 0; k < 16; k++) data[base + (size_t)k * 256] = canon_fp(x[k], fm.q, fm.qinv);
            } else {
                u64 x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = data[base + (size_t)k * 256];
                inv16<true>(x, 1u, W, Ws, q, mc.ninv, mc.ninv_s, mc.w1n, mc.w1n_s);
#pragma unroll
                for (int k = 0; k < 16; k++) data[base + (size_t)k * 256] = canon2(x[k], q);
            }
        }
    }
}
template <int LOGR>
__global__ void __launch_bounds__(kThreads, NTT_MIN_BLOCKS)
ntt_inv_passA(u64* __restrict__ data, const GRID_CONST NttJob J, NttTables T) {
    CKKS_SHARED u64 sm[LOGR == 8 ? 256 * 16 : 1];
    if (J.cnt[blockIdx.z] && blockIdx.y >= J.cnt[blockIdx.z]) return;
    const int limb = J.rows[blockIdx.z][blockIdx.y], tile = blockIdx.x;
    const int mod = J.mods[blockIdx.z][blockIdx.y];
    const size_t N = (size_t)1 << T.logn;
    data += blockIdx.z * J.dzs;
    const ModConst mc = T.mc[mod];
    if (use_fp(mc.q)) inv_passA_body<LOGR, true>(data, sm, limb, tile, N, mc, T, mod);
    else inv_passA_body<LOGR, false>(data, sm, limb, tile, N, mc, T, mod);
}

}  // namespace

void ntt_forward(const u64* src, u64* dst, const NttJob& J, const NttTables& T, dev_stream st) {
    if (J.n == 0 || J.nz == 0) return;
    const unsigned R = 1u << (T.logn - 8);
    dim3 gridB(R / 16, J.n, J.nz);
    if (T.logn == 16) {
        dim3 gridA(16, J.n, J.nz);
        LAUNCH(ntt_fwd_passA<8>, gridA, dim3(kThreads), st, src, dst, J, T);
    } else if (T.logn == 12) {
        dim3 gridA(1, J.n, J.nz);
        LAUNCH(ntt_fwd_passA<4>, gridA, dim3(kThreads), st, src, dst, J, T);
    } else {
        throw std::runtime_error("ntt: only N = 2^16 and N = 2^12 are built");
    }
    LAUNCH(ntt_fwd_passB, gridB, dim3(kThreads), st, dst, J, T);
}

void ntt_inverse(const u64* src, u64* dst, const NttJob& J, const NttTables& T, dev_stream st) {
    if (J.n == 0 || J.nz == 0) return;
    const unsigned R = 1u << (T.logn - 8);
    dim3 gridB(R / 16, J.n, J.nz);
    LAUNCH(ntt_inv_passB, gridB, dim3(kThreads), st, src, dst, J, T);
    if (T.logn == 16) {
        dim3 gridA(16, J.n, J.nz);
        LAUNCH(ntt_inv_passA<8>, gridA, dim3(kThreads), st, dst, J, T);
    } else if (T.logn == 12) {
        dim3 gridA(1, J.n, J.nz);
        LAUNCH(ntt_inv_passA<4>, gridA, dim3(kThreads), st, dst, J, T);
    } else {
        throw std::runtime_error("ntt: only N = 2^16 and N = 2^12 are built");
    }
}
