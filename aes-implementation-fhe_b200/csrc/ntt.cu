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
// for the four register rounds.  Each thread keeps 16 coefficients in registers for 4 stages
// (Harvey lazy butterflies, values in [0,4q)), with ONE shared-memory exchange per pass.
// Global accesses are 128-byte coalesced in both passes; pass B stages its stores through
// shared memory (padded 1-in-16 so the 64-bit accesses are bank-conflict free).
//
// Algorithmic bytes: 2*N*8 = 1 MiB per limb-NTT at N = 2^16; actual DRAM traffic
// is 2 MiB if the intermediate misses L2, ~1 MiB when the batch fits the 126 MB L2 (in-place).
#include "ntt.cuh"

namespace {

constexpr int kThreads = 256;
#ifndef NTT_MIN_BLOCKS
#define NTT_MIN_BLOCKS 3      // register cap 80, no spills; measured best of {2,3,4,5} (profiles/r1_ntt_occupancy_sweep.txt)
#endif

__device__ __forceinline__ int pad16(int i) { return i + (i >> 4); }

// floor(a b / 2^64) - {0,1,2}: drops the low x low partial product and the carries out of the middle column
// (3 wide multiplies instead of 4).  Good enough for a *lazy* Shoup quotient.
__device__ __forceinline__ u64 mulhi64_approx(u64 a, u64 b) {
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    const u64 t = (u64)a1 * b0, u = (u64)a0 * b1;
    return (u64)a1 * b1 + (t >> 32) + (u >> 32);
}
// x * w mod q in [0, 4q) for any 64-bit x (approximate quotient: at most 2 too small)
__device__ __forceinline__ u64 shoup_mul_lazy4(u64 x, u64 w, u64 ws, u64 q) {
    return w * x - mulhi64_approx(ws, x) * q;
}

// Moduli below 2^52 (the scale primes) take the SMALL path: no per-stage correction at all.  Forward values grow by
// at most 4q per stage (<= 65q < 2^59 after 16 stages); inverse sums double per stage and are folded back at the pass
// boundary (<= 2^9 q after 8 stages).  The 60/61-bit moduli (q_0, special primes) keep the Harvey [0,4q) / [0,2q) forms.
#ifndef NTT_VARIANT
#define NTT_VARIANT 0
#endif
#if NTT_VARIANT == 2
#define NTT_SMALL_BITS 0          // experiment: no modulus takes the SMALL path
#else
#define NTT_SMALL_BITS 52
#endif

// Forward radix-16 block rooted at table index X: 4 CT stages on x[0..15].
template <bool SMALL>
__device__ __forceinline__ void fwd16(u64 (&x)[16], u32 X, const u64* __restrict__ W, const u64* __restrict__ Ws,
                                      u64 q) {
    const u64 q2 = 2 * q, q4 = 4 * q;
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
            if (SMALL) {
                const u64 u = x[k0];
                const u64 v = shoup_mul_lazy4(x[k1], w[g], ws[g], q);
                x[k0] = u + v;
                x[k1] = u - v + q4;
            } else {
                u64 u = x[k0];
                u = u >= q2 ? u - q2 : u;
                const u64 v = shoup_mul_lazy(x[k1], w[g], ws[g], q);
                x[k0] = u + v;
                x[k1] = u - v + q2;
            }
        }
    }
}

// Inverse radix-16 block rooted at X: 4 GS stages (stage numbers stage0+1 .. stage0+4 of the pass).
// If FINAL, the last stage folds in N^-1 (scaling the sum by ninv and the twiddle by ninv).
template <bool FINAL, bool SMALL>
__device__ __forceinline__ void inv16(u64 (&x)[16], u32 X, const u64* __restrict__ W, const u64* __restrict__ Ws,
                                      u64 q, u64 ninv, u64 ninv_s, u64 w1n, u64 w1n_s, int stage0) {
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
        const u64 M = q << (stage0 + s + 2);            // SMALL: a multiple of q above every input of this stage
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int g = i / span, k0 = g * 2 * span + (i % span), k1 = k0 + span;
            const u64 u = x[k0], v = x[k1];
            if (SMALL) {
                u64 sum = u + v;
                if (FINAL && s == 3) sum = shoup_mul_lazy4(sum, ninv, ninv_s, q);
                x[k0] = sum;
                x[k1] = shoup_mul_lazy4(u - v + M, w[g], ws[g], q);
            } else {
                u64 sum = u + v;
                sum = sum >= q2 ? sum - q2 : sum;
                if (FINAL && s == 3) sum = shoup_mul_lazy(sum, ninv, ninv_s, q);
                x[k0] = sum;
                x[k1] = shoup_mul_lazy(u - v + q2, w[g], ws[g], q);
            }
        }
    }
}

// fold a lazy value below 2^7 q (forward SMALL path: <= 65q) to [0,q) by a chain of conditional subtractions: ALU-pipe
// work only, no multiplies (the Barrett variant costs five more IMAD-class instructions per value)
__device__ __forceinline__ u64 canon128(u64 v, u64 q) {
#pragma unroll
    for (int s = 6; s >= 0; s--) {
        const u64 m = q << s;
        v = v >= m ? v - m : v;
    }
    return v;
}
__device__ __forceinline__ u64 canon4(u64 v, u64 q) {
    v = v >= 2 * q ? v - 2 * q : v;
    return v >= q ? v - q : v;
}
__device__ __forceinline__ u64 canon2(u64 v, u64 q) { return v >= q ? v - q : v; }

// ---------------------------------------------------------------- forward, pass A (columns)
// LOGR = 8: R = 256 rows, tile = 16 columns x 256 rows, two radix-16 rounds (X = 1, then 16 + rr).
// LOGR = 4: R = 16 rows,  tile = 256 columns x 16 rows, one radix-16 round (X = 1).
template <int LOGR, bool SMALL>
__device__ __forceinline__ void fwd_passA_body(const u64* __restrict__ src, u64* __restrict__ dst, u64* sm, int limb,
                                               int slimb, int tile, size_t N, u64 q, const u64* __restrict__ W,
                                               const u64* __restrict__ Ws) {
    constexpr int RG = (1 << LOGR) / 16;          // row groups per column: 16 or 1
    constexpr int TC = kThreads / RG;             // columns per CTA: 16 or 256
    if (LOGR == 8) {
        FOR_THREADS {
            const int c = threadIdx.x % TC, rr = threadIdx.x / TC;
            const size_t sbase = (size_t)slimb * N + tile * TC + c;
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = src[sbase + (size_t)(rr + 16 * k) * 256];
            fwd16<SMALL>(x, 1u, W, Ws, q);
#pragma unroll
            for (int k = 0; k < 16; k++) sm[(rr + 16 * k) * TC + c] = x[k];
        }
        BLOCK_SYNC;
        FOR_THREADS {
            const int c = threadIdx.x % TC, rr = threadIdx.x / TC;
            const size_t base = (size_t)limb * N + tile * TC + c;
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = sm[(16 * rr + k) * TC + c];
            fwd16<SMALL>(x, 16u + rr, W, Ws, q);
#pragma unroll
            for (int k = 0; k < 16; k++) dst[base + (size_t)(16 * rr + k) * 256] = x[k];   // lazy
        }
    } else {
        FOR_THREADS {
            const size_t base = (size_t)limb * N + threadIdx.x, sbase = (size_t)slimb * N + threadIdx.x;
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = src[sbase + (size_t)k * 256];
            fwd16<SMALL>(x, 1u, W, Ws, q);
#pragma unroll
            for (int k = 0; k < 16; k++) dst[base + (size_t)k * 256] = x[k];
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
    const u64 q = T.mc[mod].q;
    const u64* W = T.fwd + (size_t)mod * N;
    const u64* Ws = T.fwd_s + (size_t)mod * N;
    if (q >> NTT_SMALL_BITS) fwd_passA_body<LOGR, false>(src, dst, sm, limb, slimb, tile, N, q, W, Ws);
    else fwd_passA_body<LOGR, true>(src, dst, sm, limb, slimb, tile, N, q, W, Ws);
}

// ---------------------------------------------------------------- forward, pass B (rows)
// 16 rows of 256 per CTA; row with global index Rg is rooted at table index R + Rg.
template <bool SMALL>
__device__ __forceinline__ void fwd_passB_body(u64* __restrict__ g, u64* sm, int tile, u32 Rn, const ModConst& mc,
                                               const u64* __restrict__ W, const u64* __restrict__ Ws) {
    const u64 q = mc.q;
    FOR_THREADS {
        const int jj = threadIdx.x % 16, row = threadIdx.x / 16;
        const u32 R = tile * 16 + row;
        u64 x[16];
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = g[row * 256 + jj + 16 * k];
        fwd16<SMALL>(x, Rn + R, W, Ws, q);
#pragma unroll
        for (int k = 0; k < 16; k++) sm[pad16(row * 256 + jj + 16 * k)] = x[k];
    }
    BLOCK_SYNC;
    FOR_THREADS {
        const int jj = threadIdx.x % 16, row = threadIdx.x / 16;
        const u32 R = tile * 16 + row;
        u64 x[16];
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = sm[pad16(row * 256 + 16 * jj + k)];
        fwd16<SMALL>(x, 16u * (Rn + R) + jj, W, Ws, q);
        // each thread rewrites exactly the 16 slots it just read, so no barrier is needed before this store
#pragma unroll
        for (int k = 0; k < 16; k++)
#if NTT_VARIANT == 1
            sm[pad16(row * 256 + 16 * jj + k)] = SMALL ? canon128(x[k], q) : canon4(x[k], q);
#else
            sm[pad16(row * 256 + 16 * jj + k)] = SMALL ? barrett_reduce64(x[k], mc) : canon4(x[k], q);
#endif
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
    const u64* W = T.fwd + (size_t)mod * N;
    const u64* Ws = T.fwd_s + (size_t)mod * N;
    u64* g = data + (size_t)limb * N + (size_t)tile * 16 * 256;
    if (mc.q >> NTT_SMALL_BITS) fwd_passB_body<false>(g, sm, tile, Rn, mc, W, Ws);
    else fwd_passB_body<true>(g, sm, tile, Rn, mc, W, Ws);
}

// ---------------------------------------------------------------- inverse, pass B^-1 (rows)
template <bool SMALL>
__device__ __forceinline__ void inv_passB_body(const u64* __restrict__ s_in, u64* __restrict__ d_out, u64* sm, int tile,
                                               u32 Rn, const ModConst& mc, const u64* __restrict__ W,
                                               const u64* __restrict__ Ws) {
    const u64 q = mc.q;
    FOR_THREADS {
#pragma unroll
        for (int k = 0; k < 16; k++) sm[pad16(k * 256 + threadIdx.x)] = s_in[k * 256 + threadIdx.x];
    }
    BLOCK_SYNC;
    FOR_THREADS {
        const int jj = threadIdx.x % 16, row = threadIdx.x / 16;
        const u32 R = tile * 16 + row;
        u64 x[16];
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = sm[pad16(row * 256 + 16 * jj + k)];
        inv16<false, SMALL>(x, 16u * (Rn + R) + jj, W, Ws, q, 0, 0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < 16; k++) sm[pad16(row * 256 + 16 * jj + k)] = x[k];
    }
    BLOCK_SYNC;
    FOR_THREADS {
        const int jj = threadIdx.x % 16, row = threadIdx.x / 16;
        const u32 R = tile * 16 + row;
        u64 x[16];
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = sm[pad16(row * 256 + jj + 16 * k)];
        inv16<false, SMALL>(x, Rn + R, W, Ws, q, 0, 0, 0, 0, 4);
        // SMALL: sums have grown to < 2^9 q: fold back to [0,q) at the pass boundary; else lazy [0,2q)
#pragma unroll
        for (int k = 0; k < 16; k++) d_out[row * 256 + jj + 16 * k] = SMALL ? barrett_reduce64(x[k], mc) : x[k];
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
    const u64* W = T.inv + (size_t)mod * N;
    const u64* Ws = T.inv_s + (size_t)mod * N;
    const u64* s_in = src + (size_t)slimb * N + (size_t)tile * 16 * 256;
    u64* d_out = dst + (size_t)limb * N + (size_t)tile * 16 * 256;
    if (mc.q >> NTT_SMALL_BITS) inv_passB_body<false>(s_in, d_out, sm, tile, Rn, mc, W, Ws);
    else inv_passB_body<true>(s_in, d_out, sm, tile, Rn, mc, W, Ws);
}

// ---------------------------------------------------------------- inverse, pass A^-1 (columns)
template <int LOGR, bool SMALL>
__device__ __forceinline__ void inv_passA_body(u64* __restrict__ data, u64* sm, int limb, int tile, size_t N,
                                               const ModConst& mc, const u64* __restrict__ W,
                                               const u64* __restrict__ Ws) {
    constexpr int RG = (1 << LOGR) / 16;
    constexpr int TC = kThreads / RG;
    const u64 q = mc.q;
    if (LOGR == 8) {
        FOR_THREADS {
            const int c = threadIdx.x % TC, rr = threadIdx.x / TC;
            const size_t base = (size_t)limb * N + tile * TC + c;
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = data[base + (size_t)(16 * rr + k) * 256];
            inv16<false, SMALL>(x, 16u + rr, W, Ws, q, 0, 0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < 16; k++) sm[(16 * rr + k) * TC + c] = x[k];
        }
        BLOCK_SYNC;
        FOR_THREADS {
            const int c = threadIdx.x % TC, rr = threadIdx.x / TC;
            const size_t base = (size_t)limb * N + tile * TC + c;
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = sm[(rr + 16 * k) * TC + c];
            inv16<true, SMALL>(x, 1u, W, Ws, q, mc.ninv, mc.ninv_s, mc.w1n, mc.w1n_s, 4);
#pragma unroll
            for (int k = 0; k < 16; k++)
                data[base + (size_t)(rr + 16 * k) * 256] = SMALL ? canon4(x[k], q) : canon2(x[k], q);
        }
    } else {
        FOR_THREADS {
            const size_t base = (size_t)limb * N + threadIdx.x;
            u64 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = data[base + (size_t)k * 256];
            inv16<true, SMALL>(x, 1u, W, Ws, q, mc.ninv, mc.ninv_s, mc.w1n, mc.w1n_s, 0);
#pragma unroll
            for (int k = 0; k < 16; k++) data[base + (size_t)k * 256] = SMALL ? canon4(x[k], q) : canon2(x[k], q);
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
    const u64* W = T.inv + (size_t)mod * N;
    const u64* Ws = T.inv_s + (size_t)mod * N;
    if (mc.q >> NTT_SMALL_BITS) inv_passA_body<LOGR, false>(data, sm, limb, tile, N, mc, W, Ws);
    else inv_passA_body<LOGR, true>(data, sm, limb, tile, N, mc, W, Ws);
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
