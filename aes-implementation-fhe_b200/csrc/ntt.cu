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
// for the four register rounds.  Each thread keeps 16 coefficients in registers for 4 stages.
//
// Memory staging.  Pass A needs table entries 1..255 only (the same for every CTA of a limb): its data tile and those
// twiddles are copied with cp.async (LDGSTS) in the first instructions of the kernel and the arithmetic runs from shared
// memory (no global latency inside the butterfly stages; 39 -> 34.5 us on 126 limbs).  Pass B needs 15 twiddle pairs
// per sixteen-element group, 61 KB per tile: staging them too was measured (100 KB of shared memory per CTA, two CTAs
// per SM: 54 -> 57 us on 126 limbs, and slower inside the AES step where other lanes' kernels share the SMs), so pass
// B reads its twiddles from the global tables, a pair per 128-bit load (profiles/README.md, NTT v2).
//
// Two arithmetic paths, chosen per limb (a CTA works on one limb, so the choice is uniform):
//   * integer path (the 60/61-bit moduli q_0 and the special primes): Harvey lazy butterflies with Shoup
//     twiddles, values in [0,4q) forward / [0,2q) inverse.  One Shoup multiplication costs a 64x64 high
//     product: measured 3.86 modmul/clk/SM on the B200 (profiles/r1_pipe_peaks.txt).
//   * FP64 path (the ~2^50 scale primes, q < 1.4 * 2^50): coefficients are held as exact integers in doubles and
//     multiplied with two DMUL, two DFMA, one rounding and one add (error-free product + quotient estimate):
//     measured 7.70 modmul/clk/SM, on the FP64 pipe, which the integer path leaves idle.  Values stay two-sided
//     lazy (|x| < 2^53) and are folded to |x| <= q/2 once per radix-16 block.  Every operation is exact, so the
//     canonical output is bit-identical to the integer path and to the oracle.
//
// Algorithmic bytes: 2*N*8 = 1 MiB per limb-NTT at N = 2^16.
#include "ntt.cuh"
#ifndef CKKS_EMU
#include <cooperative_groups.h>
#endif

namespace {

constexpr int kThreads = 256;
#ifndef NTT_MIN_BLOCKS_A
#define NTT_MIN_BLOCKS_A 4    // pass A: 37 KB static shared memory, register cap 64 (A/B in graph mode: profiles/README.md)
#endif
#ifndef NTT_MIN_BLOCKS_B
#define NTT_MIN_BLOCKS_B 4    // pass B: 35 KB static shared memory, register cap 64
#endif
#ifndef NTT_FP64
#define NTT_FP64 1            // 0: every limb takes the integer path (A/B measurement)
#endif
// FP64 path bound: |x| < 5.1 q must stay below 2^53 (CKKS_FP_LIMIT, common.cuh)
#define NTT_FP_LIMIT CKKS_FP_LIMIT

__device__ __forceinline__ int pad16(int i) { return i + (i >> 4); }
__device__ __forceinline__ bool use_fp(u64 q) { return NTT_FP64 && q < NTT_FP_LIMIT; }
// lazy doubles travel between the two passes (and through shared memory) in u64 words as raw bits
__device__ __forceinline__ u64 d2bits(double x) { union { double d; u64 u; } c; c.d = x; return c.u; }
__device__ __forceinline__ double bits2d(u64 x) { union { double d; u64 u; } c; c.u = x; return c.d; }

struct alignas(16) u64x2 { u64 a, b; };

// Two-pass kernels: blockIdx.y = batch item * nz + slice, blockIdx.z = item of the job (limb) -- "limb-major": the CTAs of
// item 0 of EVERY slice and batch item are scheduled first.  Item 0 of a job over Q_l is the q_0 limb, whose integer-path CTAs
// run about twice as long as the FP64 ones: started first they overlap with the rest instead of forming the tail of the
// launch (NTT_LIMB_MAJOR=0: the round-2 order, z = slice-major).  The cluster kernels keep (8, item, slice).
#ifndef NTT_LIMB_MAJOR
#define NTT_LIMB_MAJOR 1
#endif
#if NTT_LIMB_MAJOR
#define BIDX_ITEM blockIdx.z
#define BIDX_SLICE blockIdx.y
#define NTT_GRID(x, n, zb) dim3((x), (zb), (n))
#else
#define BIDX_ITEM blockIdx.y
#define BIDX_SLICE blockIdx.z
#define NTT_GRID(x, n, zb) dim3((x), (n), (zb))
#endif
struct ZB { unsigned z, b; };
__device__ __forceinline__ ZB zb_split_of(const NttJob& J, unsigned s) {
    ZB r;
    r.b = s / (unsigned)J.nz;
    r.z = s - r.b * (unsigned)J.nz;
    return r;
}
__device__ __forceinline__ ZB zb_split(const NttJob& J) { return zb_split_of(J, BIDX_SLICE); }
__device__ __forceinline__ ZB zb_split_cluster(const NttJob& J) { return zb_split_of(J, blockIdx.z); }

// ============================================================================================ twiddle accessors
// A radix-16 block rooted at table index X needs, for r = 0..3, the 2^r entries (X << r) + g.  get(r, g) returns the
// twiddle and its companion (Shoup word / quotient by q) as raw 64-bit words; get2 returns the pair g = 2 gh, 2 gh + 1.
//
// TwLin: a linear shared-memory copy.  MODE 0: entries 1..255 stored at [e - 1], block X = 1;  MODE 1: the same copy,
// blocks X = 16 + xl;  MODE 2: the four ranges of 16 consecutive blocks stored back to back (16, 32, 64, 128 entries).
template <int MODE>
__device__ __forceinline__ int tw_base(int r) {
    return MODE == 0 ? (1 << r) - 1 : MODE == 1 ? (16 << r) - 1 : (16 << r) - 16;
}
// NOC (the FP64 path): no companion word exists -- modmul_fp takes its quotient estimate from the product itself.
template <int MODE, bool NOC>
struct TwLin {
    const u64* W;
    const u64* C;
    u32 xl;
    DEV_MEMBER void get(int r, int g, u64& w, u64& c) const {
        const int i = tw_base<MODE>(r) + (int)(xl << r) + g;
        w = W[i];
        c = NOC ? 0 : C[i];
    }
    DEV_MEMBER void get2(int r, int gh, u64& w0, u64& c0, u64& w1, u64& c1) const {
        get(r, 2 * gh, w0, c0);
        get(r, 2 * gh + 1, w1, c1);
    }
};
// TwGlobal: straight from the global table (read-only path), pairs as one 128-bit load.  Used by pass B, whose
// per-thread twiddles (30 KB per tile on the FP64 path, 61 KB with Shoup companions) would cost resident CTAs if they
// were staged.
template <bool NOC>
struct TwGlobal {
    const u64* W;
    const u64* C;
    u32 X;
    DEV_MEMBER void get(int r, int g, u64& w, u64& c) const {
        w = ldg(W + (X << r) + g);
        c = NOC ? 0 : ldg(C + (X << r) + g);
    }
    DEV_MEMBER void get2(int r, int gh, u64& w0, u64& c0, u64& w1, u64& c1) const {
        ldg_pair(W + (X << r) + 2 * gh, w0, w1);
        if (NOC) { c0 = c1 = 0; return; }
        ldg_pair(C + (X << r) + 2 * gh, c0, c1);
    }
};
// TwRegs: the 15 twiddles of one radix-16 block held in registers, loaded ahead of the barrier that precedes the round
// which uses them (FP64 path of pass B: the per-thread twiddles of the second round come from L2; requested after the first
// round's stores they arrive during the barrier and the shared-memory exchange instead of stalling the first butterflies).
#ifndef NTT_TW_PRELOAD
#define NTT_TW_PRELOAD 1
#endif
struct TwRegs {
    const u64* w;
    DEV_MEMBER void get(int r, int g, u64& wo, u64& c) const { wo = w[(1 << r) - 1 + g]; c = 0; }
    DEV_MEMBER void get2(int r, int gh, u64& w0, u64& c0, u64& w1, u64& c1) const {
        w0 = w[(1 << r) - 1 + 2 * gh];
        w1 = w[(1 << r) - 1 + 2 * gh + 1];
        c0 = c1 = 0;
    }
};
__device__ __forceinline__ void tw_preload(const TwGlobal<true>& t, u64 (&w)[15]) {
    u64 c0, c1;
    t.get(0, 0, w[0], c0);
#pragma unroll
    for (int r = 1; r < 4; r++)
#pragma unroll
        for (int gh = 0; gh < (1 << (r - 1)); gh++) t.get2(r, gh, w[(1 << r) - 1 + 2 * gh], c0, w[(1 << r) - 1 + 2 * gh + 1], c1);
}
// per-thread storage that outlives a barrier: a plain local array under nvcc, one row per emulated thread otherwise
#ifdef CKKS_EMU
#define PER_THREAD_ROWS kThreads
#define PER_THREAD_ROW threadIdx.x
#else
#define PER_THREAD_ROWS 1
#define PER_THREAD_ROW 0
#endif
#ifndef NTT_PREFETCH
#define NTT_PREFETCH 0        // 0: off (default), 1: towards L1, 2: towards L2 -- measured slower on the B200 (profiles/r1_ab_prefetch.txt)
#endif
// the 1 + 2 + 4 + 8 entries (and companions) of the radix-16 block rooted at X: asked for while the first round computes
__device__ __forceinline__ void tw_prefetch(const u64* W, const u64* C, u32 X) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
        if (NTT_PREFETCH == 1) { prefetch_l1(W + (X << r)); prefetch_l1(C + (X << r)); }
        if (NTT_PREFETCH == 2) { prefetch_l2(W + (X << r)); prefetch_l2(C + (X << r)); }
    }
}
template <typename TW>
__device__ __forceinline__ void tw_stage(const TW& tw, int r, u64 (&w)[8], u64 (&c)[8]) {
    if (r == 0) tw.get(0, 0, w[0], c[0]);
#pragma unroll
    for (int gh = 0; gh < 4; gh++)
        if (r > 0 && 2 * gh < (1 << r)) tw.get2(r, gh, w[2 * gh], c[2 * gh], w[2 * gh + 1], c[2 * gh + 1]);
}

constexpr int kPassBData = 16 * 256 + 16 * 16;                   // pass B tile in the pad16 layout

// ============================================================================================ integer path
// Forward radix-16 block: 4 CT stages on x[0..15], lazy in [0,4q).
template <typename TW>
__device__ __forceinline__ void fwd16(u64 (&x)[16], const TW& tw, u64 q) {
    const u64 q2 = 2 * q;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int span = 8 >> s;
        u64 w[8], ws[8];
        tw_stage(tw, s, w, ws);
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

// Inverse radix-16 block: 4 GS stages, values kept in [0,2q).
// If FINAL, the last stage folds in N^-1 (scaling the sum by ninv and the twiddle by ninv).
template <bool FINAL, typename TW>
__device__ __forceinline__ void inv16(u64 (&x)[16], const TW& tw, u64 q, u64 ninv, u64 ninv_s, u64 w1n, u64 w1n_s) {
    const u64 q2 = 2 * q;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int span = 1 << s;
        u64 w[8], ws[8];
        if (FINAL && s == 3) { w[0] = w1n; ws[0] = w1n_s; }
        else tw_stage(tw, 3 - s, w, ws);
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
// modmul_fp / fold_fp / canon_fp: common.cuh (shared with the basis conversion)
struct FpMod {
    double q, qinv, ninv, w1n;
};

// Forward radix-16 block, 4 CT stages.  In: |x| <= 0.51 q.  A product of an input |a| = m q comes back with
// |r| <= (0.5 + m q 2^-52) q (the quotient estimate is off by |a| 2^-52 at most), so the lazy values grow as
// 0.51 -> 1.2 -> 2.2 -> 3.4 -> 5.02 q at q = 1.4 * 2^50 (tools/ntt_fp_bounds.py): below 2^53 = 5.71 q without any fold
// inside the block.  The caller folds once per block.
template <typename TW>
__device__ __forceinline__ void fwd16_fp(double (&x)[16], const TW& tw, const FpMod& m) {
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int span = 8 >> s;
        u64 w[8], wq[8];
        tw_stage(tw, s, w, wq);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int g = i / span, k0 = g * 2 * span + (i % span), k1 = k0 + span;
            const double u = x[k0];
            const double v = modmul_fp(x[k1], bits2d(w[g]), m.q, m.qinv);
            x[k0] = fadd_rn(u, v);
            x[k1] = fsub_rn(u, v);
        }
    }
}
// Inverse radix-16 block, 4 GS stages.  In: |x| <= 0.51 q.  Only the pure-sum paths double per stage: before the last
// stage x[0], x[8] hold up to 4.08 q and x[1], x[9] up to 3.5 q, everything else is a product or a sum of products.
// Folding those four keeps every intermediate below 4.9 q < 2^53 at q = 1.4 * 2^50 (tools/ntt_fp_bounds.py).
template <bool FINAL, typename TW>
__device__ __forceinline__ void inv16_fp(double (&x)[16], const TW& tw, const FpMod& m) {
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int span = 1 << s;
        if (s == 3) {
            x[0] = fold_fp(x[0], m.q, m.qinv); x[8] = fold_fp(x[8], m.q, m.qinv);
            x[1] = fold_fp(x[1], m.q, m.qinv); x[9] = fold_fp(x[9], m.q, m.qinv);
        }
        u64 w[8], wq[8];
        if (FINAL && s == 3) w[0] = d2bits(m.w1n);
        else tw_stage(tw, 3 - s, w, wq);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int g = i / span, k0 = g * 2 * span + (i % span), k1 = k0 + span;
            const double u = x[k0], v = x[k1];
            double sum = fadd_rn(u, v);
            if (FINAL && s == 3) sum = modmul_fp(sum, m.ninv, m.q, m.qinv);
            x[k0] = sum;
            x[k1] = modmul_fp(fsub_rn(u, v), bits2d(w[g]), m.q, m.qinv);
        }
    }
}

__device__ __forceinline__ FpMod fp_mod(const ModConst& mc) {
    FpMod m;
    m.q = ull2d_rn(mc.q);
    m.qinv = fdiv_rn(1.0, m.q);
    m.ninv = ull2d_rn(mc.ninv);
    m.w1n = ull2d_rn(mc.w1n);
    return m;
}

// ============================================================================================ forward, pass A
// LOGR = 8: R = 256 rows, tile = 16 columns x 256 rows, two radix-16 rounds (X = 1, then 16 + rr).
// LOGR = 4: R = 16 rows,  tile = 256 columns x 16 rows, one radix-16 round (X = 1).
// Shared memory: 4096 data words ([row][16 columns] / [k][256 columns]) + table entries 1..255 and their companions.
constexpr int kPassAWords = 4096 + 2 * 256;
// rescale lift of one coefficient x mod q_l into the row modulus (spec S6), see NttFuse
__device__ __forceinline__ u64 pro_lift(u64 x, u64 ql, const ModConst& mc) {
    const u64 h = ql >> 1;
    u64 t = x + h;
    t = t >= ql ? t - ql : t;
    // (x + h) mod q_l and h itself are below q_l: when q_l < 2 q (every pair of ~2^50 scale primes, and q_0 above them)
    // their residues modulo q are one conditional subtraction away -- no Barrett reduction per coefficient
    if (ql < 2 * mc.q) {
        const u64 r = t >= mc.q ? t - mc.q : t, hq = h >= mc.q ? h - mc.q : h;
        return sub_mod(r, hq, mc.q);
    }
    return sub_mod(barrett_reduce64(t, mc), barrett_reduce64(h, mc), mc.q);
}
// The two radix-16 rounds of pass A at N = 2^16 on a tile that already sits in shared memory (sd: [256 rows][16 columns],
// tw / tc: table entries 1..255 and their companions); shared by the one-tile kernels and the pipelined persistent one.
template <bool FP, bool PRO>
__device__ __forceinline__ void fwd_passA8_compute(u64* sd, const u64* tw, const u64* tc, u64* __restrict__ dst, int limb,
                                                   int tile, size_t N, const ModConst& mc, u64 ql,
                                                   const u64* __restrict__ gsrc = nullptr) {
    constexpr int TC = 16;
    const u64 q = mc.q;
    const FpMod fm = fp_mod(mc);
    FOR_THREADS {
        const int c = threadIdx.x % TC, rr = threadIdx.x / TC;
        const TwLin<0, FP> t1{tw, tc, 0u};
        if (FP) {
            double x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const u64 v = gsrc ? gsrc[(size_t)(rr + 16 * k) * 256 + c] : sd[(rr + 16 * k) * TC + c];
                x[k] = ull2d_rn(PRO ? pro_lift(v, ql, mc) : v);
            }
            fwd16_fp(x, t1, fm);
#pragma unroll
            for (int k = 0; k < 16; k++) sd[(rr + 16 * k) * TC + c] = d2bits(fold_fp(x[k], fm.q, fm.qinv));
        } else {
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const u64 v = gsrc ? gsrc[(size_t)(rr + 16 * k) * 256 + c] : sd[(rr + 16 * k) * TC + c];
                x[k] = PRO ? pro_lift(v, ql, mc) : v;
            }
            fwd16(x, t1, q);
#pragma unroll
            for (int k = 0; k < 16; k++) sd[(rr + 16 * k) * TC + c] = x[k];
        }
    }
    BLOCK_SYNC;
    FOR_THREADS {
        const int c = threadIdx.x % TC, rr = threadIdx.x / TC;
        const TwLin<1, FP> t2{tw, tc, (u32)rr};
        u64* d0 = dst + (size_t)limb * N + tile * TC + c;
        if (FP) {
            double x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = bits2d(sd[(16 * rr + k) * TC + c]);
            fwd16_fp(x, t2, fm);
#pragma unroll
#ifndef NTT_A_BULKST
#define NTT_A_BULKST 0        // 1: the transformed tile leaves through shared memory and 256 bulk stores of one row (A/B)
#endif
            for (int k = 0; k < 16; k++) {
                const u64 v = d2bits(fold_fp(x[k], fm.q, fm.qinv));
                if (NTT_A_BULKST) sd[(16 * rr + k) * TC + c] = v;      // the slots this thread has just read
                else d0[(size_t)(16 * rr + k) * 256] = v;
            }
        } else {
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = sd[(16 * rr + k) * TC + c];
            fwd16(x, t2, q);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (NTT_A_BULKST) sd[(16 * rr + k) * TC + c] = x[k];
                else d0[(size_t)(16 * rr + k) * 256] = x[k];   // lazy [0,4q)
            }
        }
    }
    if (NTT_A_BULKST) {
        FOR_THREADS { fence_proxy_async(); }
        BLOCK_SYNC;
        FOR_THREADS {
            bulk_s2g(dst + (size_t)limb * N + tile * TC + (size_t)threadIdx.x * 256, sd + threadIdx.x * TC, 128);
            bulk_store_commit_wait_read();
        }
    }
}
template <int LOGR, bool FP, bool PRO>
__device__ __forceinline__ void fwd_passA_body(const u64* __restrict__ src, u64* __restrict__ dst, u64* sm, int limb,
                                               int slimb, int tile, size_t N, const ModConst& mc, const NttTables& T,
                                               int mod, u64 ql) {
    constexpr int RG = (1 << LOGR) / 16;          // row groups per column: 16 or 1
    constexpr int TC = kThreads / RG;             // columns per CTA: 16 or 256
    constexpr int NTW = LOGR == 8 ? 255 : 15;
    const u64 q = mc.q;
    const u64* W = (FP ? reinterpret_cast<const u64*>(T.fwd_d) : T.fwd) + (size_t)mod * N;
    const u64* C = FP ? nullptr : T.fwd_s + (size_t)mod * N;
    const FpMod fm = fp_mod(mc);
    u64* sd = sm;
    u64* tw = sm + 4096;
    u64* tc = tw + 256;
#ifndef NTT_A_BULK
#define NTT_A_BULK 0          // 1: the tile arrives by 256 bulk copies of one 128-byte row each (A/B)
#endif
    if (LOGR == 8 && NTT_A_BULK) {
        // the 17th block of the buffer (sm + 4096 + 512 ..) holds the mbarrier; table entries 0..255 land at tw - 1 + ..
        // so that the accessors' entry e - 1 convention holds: the twiddle area is used as [e], the caller passes tw + 1
        CKKS_SHARED __align__(16) u64 s_bar[2];
        FOR_THREADS {
            if (threadIdx.x == 0) mbar_init(s_bar, 1);
        }
        BLOCK_SYNC;
        FOR_THREADS {
            const int tid = threadIdx.x;
            if (tid == 0) {
                mbar_expect_tx(s_bar, 256 * 128 + 2048 + (FP ? 0 : 2048));
                bulk_g2s(tw, W, 2048, s_bar);
                if (!FP) bulk_g2s(tc, C, 2048, s_bar);
            }
            bulk_g2s(sd + tid * 16, src + (size_t)slimb * N + tile * 16 + (size_t)tid * 256, 128, s_bar);
            mbar_wait(s_bar, 0);
        }
        BLOCK_SYNC;
        fwd_passA8_compute<FP, PRO>(sd, tw + 1, tc + 1, dst, limb, tile, N, mc, ql);
        return;
    }
    FOR_THREADS {
        const int tid = threadIdx.x;
        if (tid < NTW) { cp_async8(tw + tid, W + 1 + tid); if (!FP) cp_async8(tc + tid, C + 1 + tid); }
        if (LOGR == 8) {
            const int c = tid % TC, rr = tid / TC;
#ifndef NTT_A_DIRECT
#define NTT_A_DIRECT 0        // 1: the first round reads its tile straight from global memory (A/B)
#endif
            const u64* s0 = src + (size_t)slimb * N + tile * TC + c;
            if (!NTT_A_DIRECT) {
#pragma unroll
                for (int k = 0; k < 16; k++) cp_async8(sd + (rr + 16 * k) * TC + c, s0 + (size_t)(rr + 16 * k) * 256);
            }
        } else {
            const u64* s0 = src + (size_t)slimb * N + tid;
#pragma unroll
            for (int k = 0; k < 16; k++) cp_async8(sd + k * 256 + tid, s0 + (size_t)k * 256);
        }
        cp_async_wait_all();
    }
    BLOCK_SYNC;
    if (LOGR == 8) {
        fwd_passA8_compute<FP, PRO>(sd, tw, tc, dst, limb, tile, N, mc, ql,
                                    NTT_A_DIRECT ? src + (size_t)slimb * N + tile * TC : nullptr);
    } else {
        FOR_THREADS {
            const int tid = threadIdx.x;
            const TwLin<0, FP> t1{tw, tc, 0u};
            u64* d0 = dst + (size_t)limb * N + tid;
            if (FP) {
                double x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const u64 v = sd[k * 256 + tid];
                    x[k] = ull2d_rn(PRO ? pro_lift(v, ql, mc) : v);
                }
                fwd16_fp(x, t1, fm);
#pragma unroll
                for (int k = 0; k < 16; k++) d0[(size_t)k * 256] = d2bits(fold_fp(x[k], fm.q, fm.qinv));
            } else {
                u64 x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const u64 v = sd[k * 256 + tid];
                    x[k] = PRO ? pro_lift(v, ql, mc) : v;
                }
                fwd16(x, t1, q);
#pragma unroll
                for (int k = 0; k < 16; k++) d0[(size_t)k * 256] = x[k];
            }
        }
    }
}
template <int LOGR>
__global__ void __launch_bounds__(kThreads, NTT_MIN_BLOCKS_A)
ntt_fwd_passA(const u64* __restrict__ src, u64* __restrict__ dst, const GRID_CONST NttJob J, NttTables T) {
    CKKS_SHARED __align__(16) u64 sm[kPassAWords];
    const ZB zb = zb_split(J);
    if (J.cnt[zb.z] && BIDX_ITEM >= J.cnt[zb.z]) return;
    const int limb = J.rows[zb.z][BIDX_ITEM], slimb = J.srows[zb.z][BIDX_ITEM], tile = blockIdx.x;
    const int mod = J.mods[zb.z][BIDX_ITEM];
    const size_t N = (size_t)1 << T.logn;
    src += zb.z * J.szs + zb.b * J.sbs;
    dst += zb.z * J.dzs + zb.b * J.dbs;
    const ModConst mc = T.mc[mod];
    if (use_fp(mc.q)) fwd_passA_body<LOGR, true, false>(src, dst, sm, limb, slimb, tile, N, mc, T, mod, 0);
    else fwd_passA_body<LOGR, false, false>(src, dst, sm, limb, slimb, tile, N, mc, T, mod, 0);
}
// pass A with the rescale-lift prologue (NttFuse::pro_mod)
template <int LOGR>
__global__ void __launch_bounds__(kThreads, NTT_MIN_BLOCKS_A)
ntt_fwd_passA_lift(const u64* __restrict__ src, u64* __restrict__ dst, const GRID_CONST NttJob J, NttTables T, int pro_mod) {
    CKKS_SHARED __align__(16) u64 sm[kPassAWords];
    const ZB zb = zb_split(J);
    if (J.cnt[zb.z] && BIDX_ITEM >= J.cnt[zb.z]) return;
    const int limb = J.rows[zb.z][BIDX_ITEM], slimb = J.srows[zb.z][BIDX_ITEM], tile = blockIdx.x;
    const int mod = J.mods[zb.z][BIDX_ITEM];
    const size_t N = (size_t)1 << T.logn;
    src += zb.z * J.szs + zb.b * J.sbs;
    dst += zb.z * J.dzs + zb.b * J.dbs;
    const ModConst mc = T.mc[mod];
    const u64 ql = T.mc[pro_mod].q;
    if (use_fp(mc.q)) fwd_passA_body<LOGR, true, true>(src, dst, sm, limb, slimb, tile, N, mc, T, mod, ql);
    else fwd_passA_body<LOGR, false, true>(src, dst, sm, limb, slimb, tile, N, mc, T, mod, ql);
}

// ---------------------------------------------------------------------------------- pass A, persistent and pipelined
// One launch of at most NTT_PIPE_BLOCKS CTAs per SM; a CTA walks the tiles t = blockIdx.x, + gridDim.x, ... of the whole
// job (t = tile + 16 * (item + n * (slice + nz * batch))) with two shared-memory stages: while the two radix-16 rounds run
// on one tile, the cp.async copies of the next tile (data and the 255 twiddles of ITS modulus) are in flight into the
// other stage, so the load latency of a tile hides behind arithmetic of the SAME CTA instead of relying on the other
// resident CTAs being in a different phase (the one-tile kernels run their load and arithmetic phases in lock-step
// across a wave: ncu shows the FP64 pipe at 41-51 %).  Same arithmetic per tile: bit-identical results.
#ifndef NTT_PIPE
#define NTT_PIPE 0                 // 1: the persistent pipelined pass A (measured slower: DESIGN.md 8.2); 0: one-tile kernels
#endif
#ifndef NTT_PIPE_BLOCKS
#define NTT_PIPE_BLOCKS 3          // 2 stages x 36 KB = 72 KB per CTA
#endif
constexpr int kPipeWords = 2 * kPassAWords;
struct PassATile {
    int limb, slimb, mod, tile;
    bool valid;
    const u64* src;
    u64* dst;
};
__device__ __forceinline__ PassATile passA_tile(int t, const NttJob& J, const u64* src, u64* dst) {
    PassATile P;
    P.tile = t & 15;
    const int rest = t >> 4;
    const int y = rest % J.n;
    const unsigned zbi = (unsigned)(rest / J.n);
    const unsigned b = zbi / (unsigned)J.nz, z = zbi - b * (unsigned)J.nz;
    P.valid = !(J.cnt[z] && y >= J.cnt[z]);
    const int yy = P.valid ? y : 0;
    P.limb = J.rows[z][yy];
    P.slimb = J.srows[z][yy];
    P.mod = J.mods[z][yy];
    P.src = src + z * J.szs + b * J.sbs;
    P.dst = dst + z * J.dzs + b * J.dbs;
    return P;
}
// request tile P into the stage at `buf` (data [256][16], then the twiddles and, on the integer path, their companions)
__device__ __forceinline__ void passA_issue(const PassATile& P, u64* buf, const NttTables& T, size_t N, bool inverse) {
    if (!P.valid) return;
    const bool fp = use_fp(T.mc[P.mod].q);
    const u64* W = (inverse ? (fp ? reinterpret_cast<const u64*>(T.inv_d) : T.inv)
                            : (fp ? reinterpret_cast<const u64*>(T.fwd_d) : T.fwd)) + (size_t)P.mod * N;
    const u64* C = (inverse ? T.inv_s : T.fwd_s) + (size_t)P.mod * N;
    u64* tw = buf + 4096;
    u64* tc = tw + 256;
    FOR_THREADS {
        const int tid = threadIdx.x;
        if (tid < 255) { cp_async8(tw + tid, W + 1 + tid); if (!fp) cp_async8(tc + tid, C + 1 + tid); }
        const int c = tid % 16, rr = tid / 16;
        const u64* s0 = P.src + (size_t)P.slimb * N + P.tile * 16 + c;
        if (inverse) {
#pragma unroll
            for (int k = 0; k < 16; k++) cp_async8(buf + (16 * rr + k) * 16 + c, s0 + (size_t)(16 * rr + k) * 256);
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) cp_async8(buf + (rr + 16 * k) * 16 + c, s0 + (size_t)(rr + 16 * k) * 256);
        }
    }
}
template <bool PRO>
__global__ void __launch_bounds__(kThreads, NTT_PIPE_BLOCKS)
ntt_fwd_passA_pipe(const u64* __restrict__ src, u64* __restrict__ dst, const GRID_CONST NttJob J, NttTables T, int pro_mod) {
    DYN_SHARED_U64(sm, kPipeWords);
    const size_t N = (size_t)1 << T.logn;
    const int total = 16 * J.n * J.nz * J.nb;
    const u64 ql = PRO ? T.mc[pro_mod].q : 0;
    int t = blockIdx.x, stage = 0;
    if (t < total) passA_issue(passA_tile(t, J, src, dst), sm, T, N, false);
    for (; t < total; t += gridDim.x, stage ^= 1) {
        FOR_THREADS { cp_async_wait_all(); }
        BLOCK_SYNC;              // this tile has landed; every thread has left the previous tile's second round
        if (t + (int)gridDim.x < total)
            passA_issue(passA_tile(t + (int)gridDim.x, J, src, dst), sm + (stage ^ 1) * kPassAWords, T, N, false);
        const PassATile P = passA_tile(t, J, src, dst);
        if (!P.valid) continue;
        u64* sd = sm + stage * kPassAWords;
        const ModConst mc = T.mc[P.mod];
        if (use_fp(mc.q)) fwd_passA8_compute<true, PRO>(sd, sd + 4096, sd + 4096 + 256, P.dst, P.limb, P.tile, N, mc, ql);
        else fwd_passA8_compute<false, PRO>(sd, sd + 4096, sd + 4096 + 256, P.dst, P.limb, P.tile, N, mc, ql);
    }
}
// CTAs of a persistent launch: NTT_PIPE_BLOCKS per SM, never more than there are tiles
static int pipe_grid(int tiles) {
#ifdef CKKS_EMU
    const int slots = 6;
#else
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int slots = sms * NTT_PIPE_BLOCKS;
#endif
    return tiles < slots ? tiles : slots;
}

// ============================================================================================ forward, pass B
// 16 rows of 256 per CTA; row with global index Rg is rooted at table index R_n + Rg.  Data goes global -> registers,
// the twiddles come from the global tables (128-bit pair loads); one shared-memory exchange between the two rounds and
// one to make the final store coalesced.
template <bool FP, bool EP>
__device__ __forceinline__ void fwd_passB_body(u64* __restrict__ g, u64* sm, int tile, u32 Rn, const ModConst& mc,
                                               const NttTables& T, int mod, size_t N, const u64* __restrict__ ep_a,
                                               u64* __restrict__ ep_out, u64 sv, u64 svs, bool ep_k = false, u64 kv = 0,
                                               u64 kvs = 0) {
    const u64 q = mc.q;
    const u64* W = (FP ? reinterpret_cast<const u64*>(T.fwd_d) : T.fwd) + (size_t)mod * N;
    const u64* C = FP ? nullptr : T.fwd_s + (size_t)mod * N;
    const FpMod fm = fp_mod(mc);
    u64 twpre[PER_THREAD_ROWS][15];
    FOR_THREADS {
        const int jj = threadIdx.x % 16, row = threadIdx.x / 16;
        const TwGlobal<FP> t1{W, C, Rn + (u32)(tile * 16 + row)};
        if (NTT_PREFETCH) tw_prefetch(W, C, 16u * (Rn + (u32)(tile * 16 + row)) + (u32)jj);
        if (FP) {
            double x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = bits2d(g[row * 256 + jj + 16 * k]);
            fwd16_fp(x, t1, fm);
#pragma unroll
            for (int k = 0; k < 16; k++) sm[pad16(row * 256 + jj + 16 * k)] = d2bits(fold_fp(x[k], fm.q, fm.qinv));
            if (NTT_TW_PRELOAD)
                tw_preload(TwGlobal<true>{W, nullptr, 16u * (Rn + (u32)(tile * 16 + row)) + (u32)jj}, twpre[PER_THREAD_ROW]);
        } else {
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = g[row * 256 + jj + 16 * k];
            fwd16(x, t1, q);
#pragma unroll
            for (int k = 0; k < 16; k++) sm[pad16(row * 256 + jj + 16 * k)] = x[k];
        }
    }
    BLOCK_SYNC;
    FOR_THREADS {
        const int jj = threadIdx.x % 16, row = threadIdx.x / 16;
        const TwGlobal<FP> t2{W, C, 16u * (Rn + (u32)(tile * 16 + row)) + (u32)jj};
        // each thread rewrites exactly the 16 slots it just read, so no barrier is needed before this store
        if (FP) {
            double x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = bits2d(sm[pad16(row * 256 + 16 * jj + k)]);
            if (NTT_TW_PRELOAD) fwd16_fp(x, TwRegs{twpre[PER_THREAD_ROW]}, fm);
            else fwd16_fp(x, t2, fm);
#pragma unroll
            for (int k = 0; k < 16; k++) sm[pad16(row * 256 + 16 * jj + k)] = canon_fp(x[k], fm.q, fm.qinv);
        } else {
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = sm[pad16(row * 256 + 16 * jj + k)];
            fwd16(x, t2, q);
#pragma unroll
            for (int k = 0; k < 16; k++) sm[pad16(row * 256 + 16 * jj + k)] = canon4(x[k], q);
        }
    }
    BLOCK_SYNC;
    FOR_THREADS {
        if (EP) {
            // fused tail of ModDown / rescale: (a - NTT(x)) * s, the transform itself is never stored
            u64 a[16];
#pragma unroll
            for (int k = 0; k < 16; k++) a[k] = ep_a[k * 256 + threadIdx.x];
            if (ep_k) {                                         // level alignment: k a first (integer pipe, idle in the FP64 form)
#pragma unroll
                for (int k = 0; k < 16; k++) a[k] = shoup_mul(a[k], kv, kvs, q);
            }
#pragma unroll
            for (int k = 0; k < 16; k++)
                ep_out[k * 256 + threadIdx.x] = shoup_mul(sub_mod(a[k], sm[pad16(k * 256 + threadIdx.x)], q), sv, svs, q);
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) g[k * 256 + threadIdx.x] = sm[pad16(k * 256 + threadIdx.x)];
        }
    }
}
__global__ void __launch_bounds__(kThreads, NTT_MIN_BLOCKS_B)
ntt_fwd_passB(u64* __restrict__ data, const GRID_CONST NttJob J, NttTables T) {
    CKKS_SHARED u64 sm[kPassBData];
    const ZB zb = zb_split(J);
    if (J.cnt[zb.z] && BIDX_ITEM >= J.cnt[zb.z]) return;
    const int limb = J.rows[zb.z][BIDX_ITEM], tile = blockIdx.x;
    const int mod = J.mods[zb.z][BIDX_ITEM];
    const size_t N = (size_t)1 << T.logn;
    const u32 Rn = (u32)(N >> 8);
    data += zb.z * J.dzs + zb.b * J.dbs;
    const ModConst mc = T.mc[mod];
    u64* g = data + (size_t)limb * N + (size_t)tile * 16 * 256;
    if (use_fp(mc.q)) fwd_passB_body<true, false>(g, sm, tile, Rn, mc, T, mod, N, nullptr, nullptr, 0, 0);
    else fwd_passB_body<false, false>(g, sm, tile, Rn, mc, T, mod, N, nullptr, nullptr, 0, 0);
}
// pass B with the (a - NTT(x)) * s epilogue (NttFuse::ep_*)
__global__ void __launch_bounds__(kThreads, NTT_MIN_BLOCKS_B)
ntt_fwd_passB_ep(u64* __restrict__ data, const GRID_CONST NttJob J, NttTables T, const GRID_CONST NttFuse F) {
    CKKS_SHARED u64 sm[kPassBData];
    const ZB zb = zb_split(J);
    if (J.cnt[zb.z] && BIDX_ITEM >= J.cnt[zb.z]) return;
    const int limb = J.rows[zb.z][BIDX_ITEM], tile = blockIdx.x;
    const int mod = J.mods[zb.z][BIDX_ITEM];
    const size_t N = (size_t)1 << T.logn;
    const u32 Rn = (u32)(N >> 8);
    data += zb.z * J.dzs + zb.b * J.dbs;
    const ModConst mc = T.mc[mod];
    const size_t off = (size_t)limb * N + (size_t)tile * 16 * 256;
    u64* g = data + off;
    const u64* ep_a = F.ep_a + zb.z * F.ep_azs + zb.b * F.ep_abs + off;
    u64* ep_out = F.ep_out + zb.z * F.ep_ozs + zb.b * F.ep_obs + off;
    const u64 sv = F.s.v[BIDX_ITEM], svs = F.s.vs[BIDX_ITEM];
    const bool ek = F.ep_k != 0;
    const u64 kv = ek ? F.kl.v[BIDX_ITEM] : 0, kvs = ek ? F.kl.vs[BIDX_ITEM] : 0;
    if (use_fp(mc.q)) fwd_passB_body<true, true>(g, sm, tile, Rn, mc, T, mod, N, ep_a, ep_out, sv, svs, ek, kv, kvs);
    else fwd_passB_body<false, true>(g, sm, tile, Rn, mc, T, mod, N, ep_a, ep_out, sv, svs, ek, kv, kvs);
}

// ============================================================================================ inverse, pass B^-1
template <bool FP, bool MUL>
__device__ __forceinline__ void inv_passB_body(const u64* __restrict__ s_in, u64* __restrict__ d_out, u64* sm, int tile,
                                               u32 Rn, const ModConst& mc, const NttTables& T, int mod, size_t N,
                                               const u64* __restrict__ s_in2) {
    const u64 q = mc.q;
    const u64* W = (FP ? reinterpret_cast<const u64*>(T.inv_d) : T.inv) + (size_t)mod * N;
    const u64* C = FP ? nullptr : T.inv_s + (size_t)mod * N;
    const FpMod fm = fp_mod(mc);
    u64 twpre[PER_THREAD_ROWS][15];
    FOR_THREADS {
        if (FP && NTT_TW_PRELOAD)      // the first round's per-thread twiddles travel together with the data
            tw_preload(TwGlobal<true>{W, nullptr, 16u * (Rn + (u32)(tile * 16 + threadIdx.x / 16)) + (u32)(threadIdx.x % 16)},
                       twpre[PER_THREAD_ROW]);
        if (MUL) {
            // the transformed polynomial is the product of two NTT-domain rows (d2 = a1 * b1 of a ct x ct) formed on the fly
            u64 x[16], y[16];
#pragma unroll
            for (int k = 0; k < 16; k++) { x[k] = s_in[k * 256 + threadIdx.x]; y[k] = s_in2[k * 256 + threadIdx.x]; }
#pragma unroll
            for (int k = 0; k < 16; k++) sm[pad16(k * 256 + threadIdx.x)] = barrett_mul(x[k], y[k], mc);
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) cp_async8(sm + pad16(k * 256 + threadIdx.x), s_in + k * 256 + threadIdx.x);
            if (NTT_PREFETCH) tw_prefetch(W, C, 16u * (Rn + (u32)(tile * 16 + threadIdx.x / 16)) + (u32)(threadIdx.x % 16));
            cp_async_wait_all();
        }
    }
    BLOCK_SYNC;
    FOR_THREADS {
        const int jj = threadIdx.x % 16, row = threadIdx.x / 16;
        const TwGlobal<FP> t2{W, C, 16u * (Rn + (u32)(tile * 16 + row)) + (u32)jj};
        if (FP) {
            double x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = ull2d_rn(sm[pad16(row * 256 + 16 * jj + k)]);
            if (NTT_TW_PRELOAD) inv16_fp<false>(x, TwRegs{twpre[PER_THREAD_ROW]}, fm);
            else inv16_fp<false>(x, t2, fm);
#pragma unroll
            for (int k = 0; k < 16; k++) sm[pad16(row * 256 + 16 * jj + k)] = d2bits(fold_fp(x[k], fm.q, fm.qinv));
        } else {
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = sm[pad16(row * 256 + 16 * jj + k)];
            inv16<false>(x, t2, q, 0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < 16; k++) sm[pad16(row * 256 + 16 * jj + k)] = x[k];
        }
    }
    BLOCK_SYNC;
    FOR_THREADS {
        const int jj = threadIdx.x % 16, row = threadIdx.x / 16;
        const TwGlobal<FP> t1{W, C, Rn + (u32)(tile * 16 + row)};
        if (FP) {
            double x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = bits2d(sm[pad16(row * 256 + jj + 16 * k)]);
            inv16_fp<false>(x, t1, fm);
#pragma unroll
            for (int k = 0; k < 16; k++) d_out[row * 256 + jj + 16 * k] = d2bits(fold_fp(x[k], fm.q, fm.qinv));
        } else {
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = sm[pad16(row * 256 + jj + 16 * k)];
            inv16<false>(x, t1, q, 0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < 16; k++) d_out[row * 256 + jj + 16 * k] = x[k];       // lazy [0,2q)
        }
    }
}
__global__ void __launch_bounds__(kThreads, NTT_MIN_BLOCKS_B)
ntt_inv_passB(const u64* __restrict__ src, u64* __restrict__ dst, const GRID_CONST NttJob J, NttTables T) {
    CKKS_SHARED u64 sm[kPassBData];
    const ZB zb = zb_split(J);
    if (J.cnt[zb.z] && BIDX_ITEM >= J.cnt[zb.z]) return;
    const int limb = J.rows[zb.z][BIDX_ITEM], slimb = J.srows[zb.z][BIDX_ITEM], tile = blockIdx.x;
    const int mod = J.mods[zb.z][BIDX_ITEM];
    const size_t N = (size_t)1 << T.logn;
    const u32 Rn = (u32)(N >> 8);
    src += zb.z * J.szs + zb.b * J.sbs;
    dst += zb.z * J.dzs + zb.b * J.dbs;
    const ModConst mc = T.mc[mod];
    const u64* s_in = src + (size_t)slimb * N + (size_t)tile * 16 * 256;
    u64* d_out = dst + (size_t)limb * N + (size_t)tile * 16 * 256;
    if (use_fp(mc.q)) inv_passB_body<true, false>(s_in, d_out, sm, tile, Rn, mc, T, mod, N, nullptr);
    else inv_passB_body<false, false>(s_in, d_out, sm, tile, Rn, mc, T, mod, N, nullptr);
}
// inverse pass B of the product of two polynomials (rows srows[z][i] of src and of src2)
__global__ void __launch_bounds__(kThreads, NTT_MIN_BLOCKS_B)
ntt_inv_passB_mul(const u64* __restrict__ src, const u64* __restrict__ src2, u64* __restrict__ dst, const GRID_CONST NttJob J,
                  NttTables T) {
    CKKS_SHARED u64 sm[kPassBData];
    const ZB zb = zb_split(J);
    if (J.cnt[zb.z] && BIDX_ITEM >= J.cnt[zb.z]) return;
    const int limb = J.rows[zb.z][BIDX_ITEM], slimb = J.srows[zb.z][BIDX_ITEM], tile = blockIdx.x;
    const int mod = J.mods[zb.z][BIDX_ITEM];
    const size_t N = (size_t)1 << T.logn;
    const u32 Rn = (u32)(N >> 8);
    src += zb.z * J.szs + zb.b * J.sbs;
    src2 += zb.z * J.szs + zb.b * J.s2bs;
    dst += zb.z * J.dzs + zb.b * J.dbs;
    const ModConst mc = T.mc[mod];
    const size_t off = (size_t)slimb * N + (size_t)tile * 16 * 256;
    u64* d_out = dst + (size_t)limb * N + (size_t)tile * 16 * 256;
    if (use_fp(mc.q)) inv_passB_body<true, true>(src + off, d_out, sm, tile, Rn, mc, T, mod, N, src2 + off);
    else inv_passB_body<false, true>(src + off, d_out, sm, tile, Rn, mc, T, mod, N, src2 + off);
}

// ============================================================================================ inverse, pass A^-1
template <int LOGR, bool FP>
__device__ __forceinline__ void inv_passA_body(u64* __restrict__ data, u64* sm, int limb, int tile, size_t N,
                                               const ModConst& mc, const NttTables& T, int mod) {
    constexpr int RG = (1 << LOGR) / 16;
    constexpr int TC = kThreads / RG;
    constexpr int NTW = LOGR == 8 ? 255 : 15;
    const u64 q = mc.q;
    const u64* W = (FP ? reinterpret_cast<const u64*>(T.inv_d) : T.inv) + (size_t)mod * N;
    const u64* C = FP ? nullptr : T.inv_s + (size_t)mod * N;
    const FpMod fm = fp_mod(mc);
    u64* sd = sm;
    u64* tw = sm + 4096;
    u64* tc = tw + 256;
    FOR_THREADS {
        const int tid = threadIdx.x;
        if (tid < NTW) { cp_async8(tw + tid, W + 1 + tid); if (!FP) cp_async8(tc + tid, C + 1 + tid); }
        if (LOGR == 8) {
            const int c = tid % TC, rr = tid / TC;
            const u64* s0 = data + (size_t)limb * N + tile * TC + c;
#pragma unroll
            for (int k = 0; k < 16; k++) cp_async8(sd + (16 * rr + k) * TC + c, s0 + (size_t)(16 * rr + k) * 256);
        } else {
            const u64* s0 = data + (size_t)limb * N + tid;
#pragma unroll
            for (int k = 0; k < 16; k++) cp_async8(sd + k * 256 + tid, s0 + (size_t)k * 256);
        }
        cp_async_wait_all();
    }
    BLOCK_SYNC;
    if (LOGR == 8) {
        FOR_THREADS {
            const int c = threadIdx.x % TC, rr = threadIdx.x / TC;
            const TwLin<1, FP> t2{tw, tc, (u32)rr};
            if (FP) {
                double x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = bits2d(sd[(16 * rr + k) * TC + c]);
                inv16_fp<false>(x, t2, fm);
#pragma unroll
                for (int k = 0; k < 16; k++) sd[(16 * rr + k) * TC + c] = d2bits(fold_fp(x[k], fm.q, fm.qinv));
            } else {
                u64 x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = sd[(16 * rr + k) * TC + c];
                inv16<false>(x, t2, q, 0, 0, 0, 0);
#pragma unroll
                for (int k = 0; k < 16; k++) sd[(16 * rr + k) * TC + c] = x[k];
            }
        }
        BLOCK_SYNC;
        FOR_THREADS {
            const int c = threadIdx.x % TC, rr = threadIdx.x / TC;
            const TwLin<0, FP> t1{tw, tc, 0u};
            u64* d0 = data + (size_t)limb * N + tile * TC + c;
            if (FP) {
                double x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = bits2d(sd[(rr + 16 * k) * TC + c]);
                inv16_fp<true>(x, t1, fm);
#pragma unroll
                for (int k = 0; k < 16; k++) d0[(size_t)(rr + 16 * k) * 256] = canon_fp(x[k], fm.q, fm.qinv);
            } else {
                u64 x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = sd[(rr + 16 * k) * TC + c];
                inv16<true>(x, t1, q, mc.ninv, mc.ninv_s, mc.w1n, mc.w1n_s);
#pragma unroll
                for (int k = 0; k < 16; k++) d0[(size_t)(rr + 16 * k) * 256] = canon2(x[k], q);
            }
        }
    } else {
        FOR_THREADS {
            const int tid = threadIdx.x;
            const TwLin<0, FP> t1{tw, tc, 0u};
            u64* d0 = data + (size_t)limb * N + tid;
            if (FP) {
                double x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = bits2d(sd[k * 256 + tid]);
                inv16_fp<true>(x, t1, fm);
#pragma unroll
                for (int k = 0; k < 16; k++) d0[(size_t)k * 256] = canon_fp(x[k], fm.q, fm.qinv);
            } else {
                u64 x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = sd[k * 256 + tid];
                inv16<true>(x, t1, q, mc.ninv, mc.ninv_s, mc.w1n, mc.w1n_s);
#pragma unroll
                for (int k = 0; k < 16; k++) d0[(size_t)k * 256] = canon2(x[k], q);
            }
        }
    }
}
template <int LOGR>
__global__ void __launch_bounds__(kThreads, NTT_MIN_BLOCKS_A)
ntt_inv_passA(u64* __restrict__ data, const GRID_CONST NttJob J, NttTables T) {
    CKKS_SHARED __align__(16) u64 sm[kPassAWords];
    const ZB zb = zb_split(J);
    if (J.cnt[zb.z] && BIDX_ITEM >= J.cnt[zb.z]) return;
    const int limb = J.rows[zb.z][BIDX_ITEM], tile = blockIdx.x;
    const int mod = J.mods[zb.z][BIDX_ITEM];
    const size_t N = (size_t)1 << T.logn;
    data += zb.z * J.dzs + zb.b * J.dbs;
    const ModConst mc = T.mc[mod];
    if (use_fp(mc.q)) inv_passA_body<LOGR, true>(data, sm, limb, tile, N, mc, T, mod);
    else inv_passA_body<LOGR, false>(data, sm, limb, tile, N, mc, T, mod);
}

#ifndef CKKS_EMU
// ============================================================================================ forward, ONE kernel per transform
// EXPERIMENTAL (CKKS_NTT_CLUSTER=1, off by default): both passes in one launch, the limb never leaves the
// chip between them.  Bit-exact on the B200 against the oracle (keys, raw NTT, homomorphic operations of the GPU parity
// suite at N = 2^16); NOT yet timed -- the round's GPU budget ended with the run that validated it.  One thread-block cluster of 8 CTAs per limb, 512 threads per CTA.  Pass A: CTA c
// owns columns 32c..32c+31 of all 256 rows (two radix-16 rounds in shared memory, as fwd_passA_body with 32 columns).
// Exchange: the thread (column c, row group rr) then holds rows 16rr..16rr+15 of its column in registers; they all belong
// to CTA rr/2, and go to that CTA's shared memory over the cluster (st.shared::cluster through map_shared_rank), a warp
// writing 32 consecutive words per row.  Pass B: CTA r owns rows 32r..32r+31 and runs fwd_passB_body's two rounds from
// shared memory.  The same 68 KB buffer serves both passes: a cluster barrier separates the last pass-A read from the
// first remote write, a second one the last remote write from the first pass-B read.
constexpr int kClThreads = 512;
constexpr int kClData = 32 * 256 + 32 * 16;                     // pass-B tile of 32 rows in the pad16 layout (>= 8192)
constexpr int kClWords = kClData + 2 * 256;                     // + table entries 1..255 and companions for pass A
template <bool FP, bool PRO, bool EP>
__device__ __forceinline__ void fwd_cluster_body(const u64* __restrict__ src, u64* __restrict__ dst, u64* sm, int limb,
                                                 int slimb, int tile, size_t N, const ModConst& mc, const NttTables& T,
                                                 int mod, u64 ql, const u64* __restrict__ ep_a, u64* __restrict__ ep_out,
                                                 u64 sv, u64 svs, bool active) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const u64 q = mc.q;
    const u64* W = (FP ? reinterpret_cast<const u64*>(T.fwd_d) : T.fwd) + (size_t)mod * N;
    const u64* C = FP ? nullptr : T.fwd_s + (size_t)mod * N;
    const FpMod fm = fp_mod(mc);
    u64* sd = sm;
    u64* tw = sm + kClData;
    u64* tc = tw + 256;
    const int tid = threadIdx.x;
    // ---- pass A: 32 columns x 256 rows
    {
        const int c = tid % 32, rr = tid / 32;                  // rr = 0..15
        if (tid < 255) { cp_async8(tw + tid, W + 1 + tid); if (!FP) cp_async8(tc + tid, C + 1 + tid); }
        const u64* s0 = src + (size_t)slimb * N + tile * 32 + c;
#pragma unroll
        for (int k = 0; k < 16; k++) cp_async8(sd + (rr + 16 * k) * 32 + c, s0 + (size_t)(rr + 16 * k) * 256);
        cp_async_wait_all();
        __syncthreads();
        const TwLin<0, FP> t1{tw, tc, 0u};
        const TwLin<1, FP> t2{tw, tc, (u32)rr};
        u64 y[16];
        if (FP) {
            double x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const u64 v = sd[(rr + 16 * k) * 32 + c];
                x[k] = ull2d_rn(PRO ? pro_lift(v, ql, mc) : v);
            }
            fwd16_fp(x, t1, fm);
#pragma unroll
            for (int k = 0; k < 16; k++) sd[(rr + 16 * k) * 32 + c] = d2bits(fold_fp(x[k], fm.q, fm.qinv));
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = bits2d(sd[(16 * rr + k) * 32 + c]);
            fwd16_fp(x, t2, fm);
#pragma unroll
            for (int k = 0; k < 16; k++) y[k] = d2bits(fold_fp(x[k], fm.q, fm.qinv));
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const u64 v = sd[(rr + 16 * k) * 32 + c];
                y[k] = PRO ? pro_lift(v, ql, mc) : v;
            }
            fwd16(y, t1, q);
#pragma unroll
            for (int k = 0; k < 16; k++) sd[(rr + 16 * k) * 32 + c] = y[k];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; k++) y[k] = sd[(16 * rr + k) * 32 + c];
            fwd16(y, t2, q);                                    // lazy [0,4q), as between today's two passes
        }
        // ---- exchange: rows 16rr..16rr+15 of column tile*32 + c belong to CTA rr/2
        cluster.sync();                                         // every CTA has finished reading its pass-A tile
        u64* remote = cluster.map_shared_rank(sd, (unsigned)(rr >> 1));
        const int col = tile * 32 + c, lrow0 = 16 * (rr & 1);
#pragma unroll
        for (int k = 0; k < 16; k++) remote[pad16((lrow0 + k) * 256 + col)] = y[k];
        cluster.sync();                                         // every remote write has landed
    }
    // ---- pass B: 32 rows x 256 columns, from shared memory
    {
        const int jj = tid % 16, row = tid / 16;                // row = 0..31
        const u32 Rn = (u32)(N >> 8);
        const TwGlobal<FP> t1{W, C, Rn + (u32)(tile * 32 + row)};
        const TwGlobal<FP> t2{W, C, 16u * (Rn + (u32)(tile * 32 + row)) + (u32)jj};
        if (FP) {
            double x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = bits2d(sd[pad16(row * 256 + jj + 16 * k)]);
            fwd16_fp(x, t1, fm);
#pragma unroll
            for (int k = 0; k < 16; k++) sd[pad16(row * 256 + jj + 16 * k)] = d2bits(fold_fp(x[k], fm.q, fm.qinv));
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = bits2d(sd[pad16(row * 256 + 16 * jj + k)]);
            fwd16_fp(x, t2, fm);
#pragma unroll
            for (int k = 0; k < 16; k++) sd[pad16(row * 256 + 16 * jj + k)] = canon_fp(x[k], fm.q, fm.qinv);
        } else {
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = sd[pad16(row * 256 + jj + 16 * k)];
            fwd16(x, t1, q);
#pragma unroll
            for (int k = 0; k < 16; k++) sd[pad16(row * 256 + jj + 16 * k)] = x[k];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = sd[pad16(row * 256 + 16 * jj + k)];
            fwd16(x, t2, q);
#pragma unroll
            for (int k = 0; k < 16; k++) sd[pad16(row * 256 + 16 * jj + k)] = canon4(x[k], q);
        }
        __syncthreads();
        const size_t off = (size_t)limb * N + (size_t)tile * 32 * 256;
        if (!active) return;                                    // surplus cluster of a ragged batch: no global store
        if (EP) {                                               // (a - NTT(x)) * s, as ntt_fwd_passB_ep
#pragma unroll
            for (int k = 0; k < 16; k++)
                ep_out[off + k * 512 + tid] =
                    shoup_mul(sub_mod(ep_a[off + k * 512 + tid], sd[pad16(k * 512 + tid)], q), sv, svs, q);
        } else {
            u64* g = dst + off;
#pragma unroll
            for (int k = 0; k < 16; k++) g[k * 512 + tid] = sd[pad16(k * 512 + tid)];
        }
    }
}
// y index of a CTA: no early exit -- every CTA of a cluster must reach the cluster barriers.  A z-slice with fewer items than
// the launch's y extent gives its surplus clusters a clamped (repeated) item to READ and sets active = false: they run the
// barriers but store nothing (an in-place transform of a ragged batch would otherwise be transformed twice).
__device__ __forceinline__ int cluster_item(const NttJob& J, const ZB& zb, bool& active) {
    active = !(J.cnt[zb.z] && blockIdx.y >= J.cnt[zb.z]);
    return active ? (int)blockIdx.y : J.cnt[zb.z] - 1;
}
// forward transform with the rescale-lift prologue and / or the (a - NTT(x)) * s epilogue (NttFuse), one kernel
template <bool PRO, bool EP>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(kClThreads, 2)
ntt_fwd_cluster_fused(const u64* __restrict__ src, u64* __restrict__ dst, const GRID_CONST NttJob J, NttTables T,
                      const GRID_CONST NttFuse F) {
    extern __shared__ __align__(16) u64 sm_cl[];
    const ZB zb = zb_split_cluster(J);
    bool active;
    const int y = cluster_item(J, zb, active);
    const int limb = J.rows[zb.z][y], slimb = J.srows[zb.z][y], tile = blockIdx.x;
    const int mod = J.mods[zb.z][y];
    const size_t N = (size_t)1 << T.logn;
    src += zb.z * J.szs + zb.b * J.sbs;
    dst += zb.z * J.dzs + zb.b * J.dbs;
    const ModConst mc = T.mc[mod];
    const u64 ql = PRO ? T.mc[F.pro_mod].q : 0;
    const u64* ep_a = EP ? F.ep_a + zb.z * F.ep_azs + zb.b * F.ep_abs : nullptr;
    u64* ep_out = EP ? F.ep_out + zb.z * F.ep_ozs + zb.b * F.ep_obs : nullptr;
    const u64 sv = EP ? F.s.v[y] : 0, svs = EP ? F.s.vs[y] : 0;
    if (use_fp(mc.q)) fwd_cluster_body<true, PRO, EP>(src, dst, sm_cl, limb, slimb, tile, N, mc, T, mod, ql, ep_a, ep_out, sv, svs, active);
    else fwd_cluster_body<false, PRO, EP>(src, dst, sm_cl, limb, slimb, tile, N, mc, T, mod, ql, ep_a, ep_out, sv, svs, active);
}

// ============================================================================================ inverse, ONE kernel per transform
// (NOT yet run on hardware.)  The forward structure backwards: CTA r transforms rows 32r..32r+31 (pass B^-1, two radix-16
// rounds from shared memory), the thread (row, jj) then holds columns jj + 16k of its row, which belong to CTA k/2; they
// cross the cluster into that CTA's [256 rows][32 columns] tile, and pass A^-1 finishes with the 1/N scaling.
template <bool FP, bool MUL>
__device__ __forceinline__ void inv_cluster_body(const u64* __restrict__ src, const u64* __restrict__ src2,
                                                 u64* __restrict__ dst, u64* sm, int limb, int slimb, int tile, size_t N,
                                                 const ModConst& mc, const NttTables& T, int mod, bool active) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const u64 q = mc.q;
    const u64* W = (FP ? reinterpret_cast<const u64*>(T.inv_d) : T.inv) + (size_t)mod * N;
    const u64* C = FP ? nullptr : T.inv_s + (size_t)mod * N;
    const FpMod fm = fp_mod(mc);
    u64* sd = sm;
    u64* tw = sm + kClData;
    u64* tc = tw + 256;
    const int tid = threadIdx.x;
    u64 y[16];
    {
        // ---- pass B^-1: 32 rows x 256 columns
        const size_t off = (size_t)slimb * N + (size_t)tile * 32 * 256;
        if (tid < 255) { cp_async8(tw + tid, W + 1 + tid); if (!FP) cp_async8(tc + tid, C + 1 + tid); }
        if (MUL) {
#pragma unroll
            for (int k = 0; k < 16; k++)
                sd[pad16(k * 512 + tid)] = barrett_mul(src[off + k * 512 + tid], src2[off + k * 512 + tid], mc);
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) cp_async8(sd + pad16(k * 512 + tid), src + off + k * 512 + tid);
        }
        cp_async_wait_all();
        __syncthreads();
        const int jj = tid % 16, row = tid / 16;
        const u32 Rn = (u32)(N >> 8);
        const TwGlobal<FP> t2{W, C, 16u * (Rn + (u32)(tile * 32 + row)) + (u32)jj};
        const TwGlobal<FP> t1{W, C, Rn + (u32)(tile * 32 + row)};
        if (FP) {
            double x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = ull2d_rn(sd[pad16(row * 256 + 16 * jj + k)]);
            inv16_fp<false>(x, t2, fm);
#pragma unroll
            for (int k = 0; k < 16; k++) sd[pad16(row * 256 + 16 * jj + k)] = d2bits(fold_fp(x[k], fm.q, fm.qinv));
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = bits2d(sd[pad16(row * 256 + jj + 16 * k)]);
            inv16_fp<false>(x, t1, fm);
#pragma unroll
            for (int k = 0; k < 16; k++) y[k] = d2bits(fold_fp(x[k], fm.q, fm.qinv));
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) y[k] = sd[pad16(row * 256 + 16 * jj + k)];
            inv16<false>(y, t2, q, 0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < 16; k++) sd[pad16(row * 256 + 16 * jj + k)] = y[k];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; k++) y[k] = sd[pad16(row * 256 + jj + 16 * k)];
            inv16<false>(y, t1, q, 0, 0, 0, 0);                 // lazy [0,2q), as between today's two passes
        }
        // ---- exchange: column jj + 16k of global row 32 tile + row belongs to CTA k/2, local column jj + 16 (k & 1)
        cluster.sync();                                         // every CTA has finished reading its pass-B tile
        const int grow = tile * 32 + row;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            u64* remote = cluster.map_shared_rank(sd, (unsigned)(k >> 1));
            remote[grow * 32 + jj + 16 * (k & 1)] = y[k];
        }
        cluster.sync();                                         // every remote write has landed
    }
    {
        // ---- pass A^-1: 32 columns x 256 rows, 1/N folded into the last stage
        if (!active) return;                                    // surplus cluster of a ragged batch: no global store
        const int c = tid % 32, rr = tid / 32;
        const TwLin<1, FP> t2{tw, tc, (u32)rr};
        const TwLin<0, FP> t1{tw, tc, 0u};
        u64* d0 = dst + (size_t)limb * N + tile * 32 + c;
        if (FP) {
            double x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = bits2d(sd[(16 * rr + k) * 32 + c]);
            inv16_fp<false>(x, t2, fm);
#pragma unroll
            for (int k = 0; k < 16; k++) sd[(16 * rr + k) * 32 + c] = d2bits(fold_fp(x[k], fm.q, fm.qinv));
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = bits2d(sd[(rr + 16 * k) * 32 + c]);
            inv16_fp<true>(x, t1, fm);
#pragma unroll
            for (int k = 0; k < 16; k++) d0[(size_t)(rr + 16 * k) * 256] = canon_fp(x[k], fm.q, fm.qinv);
        } else {
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = sd[(16 * rr + k) * 32 + c];
            inv16<false>(x, t2, q, 0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < 16; k++) sd[(16 * rr + k) * 32 + c] = x[k];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = sd[(rr + 16 * k) * 32 + c];
            inv16<true>(x, t1, q, mc.ninv, mc.ninv_s, mc.w1n, mc.w1n_s);
#pragma unroll
            for (int k = 0; k < 16; k++) d0[(size_t)(rr + 16 * k) * 256] = canon2(x[k], q);
        }
    }
}
template <bool MUL>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(kClThreads, 2)
ntt_inv_cluster(const u64* __restrict__ src, const u64* __restrict__ src2, u64* __restrict__ dst, const GRID_CONST NttJob J,
                NttTables T) {
    extern __shared__ __align__(16) u64 sm_cl[];
    const ZB zb = zb_split_cluster(J);
    bool active;
    const int y = cluster_item(J, zb, active);
    const int limb = J.rows[zb.z][y], slimb = J.srows[zb.z][y], tile = blockIdx.x;
    const int mod = J.mods[zb.z][y];
    const size_t N = (size_t)1 << T.logn;
    src += zb.z * J.szs + zb.b * J.sbs;
    if (MUL) src2 += zb.z * J.szs + zb.b * J.s2bs;
    dst += zb.z * J.dzs + zb.b * J.dbs;
    const ModConst mc = T.mc[mod];
    if (use_fp(mc.q)) inv_cluster_body<true, MUL>(src, src2, dst, sm_cl, limb, slimb, tile, N, mc, T, mod, active);
    else inv_cluster_body<false, MUL>(src, src2, dst, sm_cl, limb, slimb, tile, N, mc, T, mod, active);
}

__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(kClThreads, 2)
ntt_fwd_cluster(const u64* __restrict__ src, u64* __restrict__ dst, const GRID_CONST NttJob J, NttTables T) {
    extern __shared__ __align__(16) u64 sm_cl[];
    const ZB zb = zb_split_cluster(J);
    bool active;
    const int y = cluster_item(J, zb, active);
    const int limb = J.rows[zb.z][y], slimb = J.srows[zb.z][y], tile = blockIdx.x;
    const int mod = J.mods[zb.z][y];
    const size_t N = (size_t)1 << T.logn;
    src += zb.z * J.szs + zb.b * J.sbs;
    dst += zb.z * J.dzs + zb.b * J.dbs;
    const ModConst mc = T.mc[mod];
    if (use_fp(mc.q)) fwd_cluster_body<true, false, false>(src, dst, sm_cl, limb, slimb, tile, N, mc, T, mod, 0, nullptr, nullptr, 0, 0, active);
    else fwd_cluster_body<false, false, false>(src, dst, sm_cl, limb, slimb, tile, N, mc, T, mod, 0, nullptr, nullptr, 0, 0, active);
}
#endif

}  // namespace

static inline int nbz(const NttJob& J) { return J.nz * (J.nb > 1 ? J.nb : 1); }
static NttJob norm(const NttJob& J) { NttJob K = J; if (K.nb < 1) K.nb = 1; return K; }

void ntt_forward(const u64* src, u64* dst, const NttJob& J0, const NttTables& T, dev_stream st) {
    if (J0.n == 0 || J0.nz == 0) return;
    const NttJob J = norm(J0);
#ifndef CKKS_EMU
    if (T.cluster >= 1 && T.logn == 16) {
        LAUNCH_DYN(ntt_fwd_cluster, dim3(8, J.n, nbz(J)), dim3(kClThreads), kClWords * sizeof(u64), st, src, dst, J, T);
        return;
    }
#endif
    const unsigned R = 1u << (T.logn - 8);
    dim3 gridB = NTT_GRID(R / 16, J.n, nbz(J));
    if (T.logn == 16) {
        dim3 gridA = NTT_GRID(16, J.n, nbz(J));
        if (NTT_PIPE)
            LAUNCH_DYN(ntt_fwd_passA_pipe<false>, dim3(pipe_grid(16 * J.n * nbz(J))), dim3(kThreads), kPipeWords * sizeof(u64), st,
                       src, dst, J, T, 0);
        else
            LAUNCH(ntt_fwd_passA<8>, gridA, dim3(kThreads), st, src, dst, J, T);
    } else if (T.logn == 12) {
        dim3 gridA = NTT_GRID(1, J.n, nbz(J));
        LAUNCH(ntt_fwd_passA<4>, gridA, dim3(kThreads), st, src, dst, J, T);
    } else {
        throw std::runtime_error("ntt: only N = 2^16 and N = 2^12 are built");
    }
    LAUNCH(ntt_fwd_passB, gridB, dim3(kThreads), st, dst, J, T);
}

void ntt_forward_fused(const u64* src, u64* dst, const NttJob& J0, const NttTables& T, const NttFuse& F, dev_stream st) {
    if (J0.n == 0 || J0.nz == 0) return;
    const NttJob J = norm(J0);
#ifndef CKKS_EMU
    if (T.cluster >= 2 && T.logn == 16) {
        if (F.ep_k) throw std::runtime_error("ntt: the cluster transform has no ep_k epilogue");
        const dim3 g(8, J.n, nbz(J)), b(kClThreads);
        const size_t smem = kClWords * sizeof(u64);
        const bool pro = F.pro_mod >= 0, ep = F.ep_out != nullptr;
        if (pro && ep) LAUNCH_DYN((ntt_fwd_cluster_fused<true, true>), g, b, smem, st, src, dst, J, T, F);
        else if (pro) LAUNCH_DYN((ntt_fwd_cluster_fused<true, false>), g, b, smem, st, src, dst, J, T, F);
        else if (ep) LAUNCH_DYN((ntt_fwd_cluster_fused<false, true>), g, b, smem, st, src, dst, J, T, F);
        else LAUNCH_DYN(ntt_fwd_cluster, g, b, smem, st, src, dst, J, T);
        return;
    }
#endif
    const unsigned R = 1u << (T.logn - 8);
    dim3 gridB = NTT_GRID(R / 16, J.n, nbz(J));
    if (T.logn != 16 && T.logn != 12) throw std::runtime_error("ntt: only N = 2^16 and N = 2^12 are built");
    dim3 gridA = NTT_GRID(T.logn == 16 ? 16 : 1, J.n, nbz(J));
    const dim3 gridP(pipe_grid(16 * J.n * nbz(J)));
    const size_t smemP = kPipeWords * sizeof(u64);
    if (F.pro_mod >= 0) {
        if (T.logn == 16 && NTT_PIPE) LAUNCH_DYN(ntt_fwd_passA_pipe<true>, gridP, dim3(kThreads), smemP, st, src, dst, J, T, F.pro_mod);
        else if (T.logn == 16) LAUNCH(ntt_fwd_passA_lift<8>, gridA, dim3(kThreads), st, src, dst, J, T, F.pro_mod);
        else LAUNCH(ntt_fwd_passA_lift<4>, gridA, dim3(kThreads), st, src, dst, J, T, F.pro_mod);
    } else {
        if (T.logn == 16 && NTT_PIPE) LAUNCH_DYN(ntt_fwd_passA_pipe<false>, gridP, dim3(kThreads), smemP, st, src, dst, J, T, 0);
        else if (T.logn == 16) LAUNCH(ntt_fwd_passA<8>, gridA, dim3(kThreads), st, src, dst, J, T);
        else LAUNCH(ntt_fwd_passA<4>, gridA, dim3(kThreads), st, src, dst, J, T);
    }
    if (F.ep_out) LAUNCH(ntt_fwd_passB_ep, gridB, dim3(kThreads), st, dst, J, T, F);
    else LAUNCH(ntt_fwd_passB, gridB, dim3(kThreads), st, dst, J, T);
}

void ntt_inverse(const u64* src, u64* dst, const NttJob& J0, const NttTables& T, dev_stream st, const u64* src2) {
    if (J0.n == 0 || J0.nz == 0) return;
    const NttJob J = norm(J0);
#ifndef CKKS_EMU
    if (T.cluster >= 2 && T.logn == 16) {
        const dim3 g(8, J.n, nbz(J)), b(kClThreads);
        const size_t smem = kClWords * sizeof(u64);
        if (src2) LAUNCH_DYN(ntt_inv_cluster<true>, g, b, smem, st, src, src2, dst, J, T);
        else LAUNCH_DYN(ntt_inv_cluster<false>, g, b, smem, st, src, src2, dst, J, T);
        return;
    }
#endif
    const unsigned R = 1u << (T.logn - 8);
    dim3 gridB = NTT_GRID(R / 16, J.n, nbz(J));
    if (src2) LAUNCH(ntt_inv_passB_mul, gridB, dim3(kThreads), st, src, src2, dst, J, T);
    else LAUNCH(ntt_inv_passB, gridB, dim3(kThreads), st, src, dst, J, T);
    if (T.logn == 16) {
        dim3 gridA = NTT_GRID(16, J.n, nbz(J));
        LAUNCH(ntt_inv_passA<8>, gridA, dim3(kThreads), st, dst, J, T);
    } else if (T.logn == 12) {
        dim3 gridA = NTT_GRID(1, J.n, nbz(J));
        LAUNCH(ntt_inv_passA<4>, gridA, dim3(kThreads), st, dst, J, T);
    } else {
        throw std::runtime_error("ntt: only N = 2^16 and N = 2^12 are built");
    }
}
