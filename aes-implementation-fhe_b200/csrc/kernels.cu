// kernels.cu -- element-wise, permutation, basis-conversion, sampling and embedding kernels.
//
// All RNS kernels work on "limb batches": row r of a [rows][N] uint64 array is a residue
// polynomial modulo mc[L.idx[r]].  Every kernel is HBM-bound streaming work except
// base_convert (integer-pipe bound: ns*(nt+1) Shoup multiplications per coefficient).
// Grids are rows x N/256 CTAs (x npoly); a 22-limb ciphertext gives 11 264 CTAs, i.e. many
// waves over 148 SMs, and all global accesses are unit-stride 8-byte per lane.
#include "kernels.cuh"

namespace {

constexpr int TPB = 256;

// ------------------------------------------------------------------ element-wise
enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2 };
// blockIdx.z selects the polynomial (and the batch item, batch-major, when ps.npoly is set); each pointer has its own
// polynomial and batch strides (elements)
struct ZOff { size_t out, a, b; };
__device__ __forceinline__ ZOff zoff(const PolyStride& ps) {
    const unsigned z = blockIdx.z;
    const unsigned bi = ps.npoly ? z / (unsigned)ps.npoly : 0u, k = ps.npoly ? z - bi * (unsigned)ps.npoly : z;
    ZOff o;
    o.out = k * ps.out + bi * ps.bout;
    o.a = k * ps.a + bi * ps.ba;
    o.b = k * ps.b + bi * ps.bb;
    return o;
}
template <int OP>
__global__ void k_binop(KShape S, u64* __restrict__ out, const u64* __restrict__ a, const u64* __restrict__ b,
                        const GRID_CONST LimbList L, PolyStride ps) {
    const int row = blockIdx.y;
    const ModConst m = S.mc[L.idx[row]];
    const ZOff o = zoff(ps);
    FOR_THREADS {
        const size_t i = ((size_t)row << S.logn) + blockIdx.x * TPB + threadIdx.x;
        const u64 x = a[i + o.a], y = b[i + o.b];
        out[i + o.out] = OP == OP_ADD ? add_mod(x, y, m.q) : OP == OP_SUB ? sub_mod(x, y, m.q) : barrett_mul(x, y, m);
    }
}
__global__ void k_neg(KShape S, u64* __restrict__ out, const u64* __restrict__ a, const GRID_CONST LimbList L, PolyStride ps) {
    const int row = blockIdx.y;
    const u64 q = S.mc[L.idx[row]].q;
    const ZOff o = zoff(ps);
    FOR_THREADS {
        const size_t i = ((size_t)row << S.logn) + blockIdx.x * TPB + threadIdx.x;
        out[i + o.out] = neg_mod(a[i + o.a], q);
    }
}
// plain copy of rows (limb drop of a batched ciphertext: one launch instead of one copy per polynomial)
__global__ void k_copy(KShape S, u64* __restrict__ out, const u64* __restrict__ a, PolyStride ps) {
    const ZOff o = zoff(ps);
    FOR_THREADS {
        const size_t i = ((size_t)blockIdx.y << S.logn) + blockIdx.x * TPB + threadIdx.x;
        out[i + o.out] = a[i + o.a];
    }
}

// out[r] = a[r] * s[r]  (per-row scalar with Shoup companion)
__global__ void k_mul_scalar(KShape S, u64* __restrict__ out, const u64* __restrict__ a, const GRID_CONST LimbList L, const GRID_CONST ScalarList Sc,
                             PolyStride ps) {
    const int row = blockIdx.y;
    const u64 q = S.mc[L.idx[row]].q;
    const ZOff o = zoff(ps);
    FOR_THREADS {
        const size_t i = ((size_t)row << S.logn) + blockIdx.x * TPB + threadIdx.x;
        out[i + o.out] = shoup_mul(a[i + o.a], Sc.v[row], Sc.vs[row], q);
    }
}

// out[r] = (a[r] - b[r]) * s[r]
__global__ void k_sub_mul_scalar(KShape S, u64* __restrict__ out, const u64* __restrict__ a,
                                 const u64* __restrict__ b, const GRID_CONST LimbList L, const GRID_CONST ScalarList Sc, PolyStride ps) {
    const int row = blockIdx.y;
    const u64 q = S.mc[L.idx[row]].q;
    const ZOff o = zoff(ps);
    FOR_THREADS {
        const size_t i = ((size_t)row << S.logn) + blockIdx.x * TPB + threadIdx.x;
        out[i + o.out] = shoup_mul(sub_mod(a[i + o.a], b[i + o.b], q), Sc.v[row], Sc.vs[row], q);
    }
}

// multiply by the constant polynomial R + I X^(N/2): first half of the bit-reversed NTT array sees cp, second half cm
template <int ACC>
__global__ void k_mul_const(KShape S, u64* __restrict__ out, const u64* __restrict__ a, const GRID_CONST LimbList L, const GRID_CONST ScalarList CP,
                            const GRID_CONST ScalarList CM, PolyStride ps) {
    const int row = blockIdx.y;
    const u64 q = S.mc[L.idx[row]].q;
    const bool lo = blockIdx.x < (gridDim.x >> 1);
    const u64 c = lo ? CP.v[row] : CM.v[row], cs = lo ? CP.vs[row] : CM.vs[row];
    const ZOff zo = zoff(ps);
    FOR_THREADS {
        const size_t i = ((size_t)row << S.logn) + blockIdx.x * TPB + threadIdx.x;
        const u64 t = shoup_mul(a[i + zo.a], c, cs, q);
        u64* o = out + i + zo.out;
        *o = ACC ? add_mod(*o, t, q) : t;
    }
}
// polynomial 0 of every batch item receives the constant, the other polynomials are copied (one launch per ciphertext)
__global__ void k_add_const(KShape S, u64* __restrict__ out, const u64* __restrict__ a, const GRID_CONST LimbList L, const GRID_CONST ScalarList CP,
                            const GRID_CONST ScalarList CM, PolyStride ps, int npoly) {
    const int row = blockIdx.y;
    const u64 q = S.mc[L.idx[row]].q;
    const bool first = (blockIdx.z % (unsigned)npoly) == 0;
    const u64 c = !first ? 0 : blockIdx.x < (gridDim.x >> 1) ? CP.v[row] : CM.v[row];
    const ZOff o = zoff(ps);
    FOR_THREADS {
        const size_t i = ((size_t)row << S.logn) + blockIdx.x * TPB + threadIdx.x;
        out[i + o.out] = add_mod(a[i + o.a], c, q);
    }
}

// tensor product of two 2-polynomial ciphertexts, nl limbs each: d0=a0 b0, d1=a0 b1+a1 b0, d2=a1 b1
__global__ void k_tensor(KShape S, u64* __restrict__ d, const u64* __restrict__ a, const u64* __restrict__ b,
                         const GRID_CONST LimbList L, int nl, PolyStride ps) {
    const int row = blockIdx.y;
    const ModConst m = S.mc[L.idx[row]];
    const size_t P = (size_t)nl << S.logn;
    d += blockIdx.z * ps.bout; a += blockIdx.z * ps.ba; b += blockIdx.z * ps.bb;      // blockIdx.z: batch item
    FOR_THREADS {
        const size_t i = ((size_t)row << S.logn) + blockIdx.x * TPB + threadIdx.x;
        const u64 a0 = a[i], a1 = a[i + P], b0 = b[i], b1 = b[i + P];
        d[i] = barrett_mul(a0, b0, m);
        u64 hi = 0, lo = 0;
        mac128(hi, lo, a0, b1);
        mac128(hi, lo, a1, b0);
        d[i + P] = barrett_reduce128(hi, lo, m);
        d[i + 2 * P] = barrett_mul(a1, b1, m);
    }
}

// out[r][k] = a[r][perm[k]]
__global__ void k_permute(KShape S, u64* __restrict__ out, const u64* __restrict__ a, const u32* __restrict__ perm,
                          PolyStride ps) {
    const ZOff o = zoff(ps);
    FOR_THREADS {
        const u32 k = blockIdx.x * TPB + threadIdx.x;
        const size_t base = (size_t)blockIdx.y << S.logn;
        out[base + o.out + k] = a[base + o.a + ldg(perm + k)];
    }
}

// key-switch inner product over `beta` digits.  ext: [beta][rows][N]; evk: [dnum][2][evk_rows][N];
// working row r uses evk row ERow.idx[r].  acc: [2][rows][N].  Row r < nq belongs to digit r / alpha: for that
// digit the value is the NTT-domain input itself (`own`, [nq][N]) -- ext never holds a copy of it.
// If perm != null the digits are read through the Galois gather (hoisted rotation: automorphism applied after
// the decomposition).
// If addend != null ([2][nq][N], NTT domain) the q rows also receive P * addend (merged relinearisation:
// acc = <digits, evk> + P (d0, d1), so that one ModDown by P q_l .. also performs the rescale, spec S6b).
// TENSOR: the switched polynomial is the d2 of a ct x ct product that was never written out -- the digit's own rows are
// a1*b1 computed here, and the addend is P*(a0 b0, a0 b1 + a1 b0) from the operands (`own` = a, `addend` = b, both
// [2][nq][N]; same formulas as k_tensor, so the result is bit-identical to tensor + inner product).
#define KS_MAX_BETA 6      /* == NTT_MAX_Z: the engine refuses parameter sets with more digits */
template <bool TENSOR>
__global__ void k_ks_inner(KShape S, u64* __restrict__ acc, const u64* __restrict__ ext, const u64* __restrict__ own,
                           const u64* __restrict__ evk, const u32* __restrict__ perm, const GRID_CONST LimbList L,
                           const GRID_CONST LimbList ERow, int beta, int rows, int evk_rows, int nq, int alpha,
                           const u64* __restrict__ addend, const GRID_CONST ScalarList PmodQ, int accumulate,
                           const GRID_CONST KsBatch kb) {
    const int row = blockIdx.y;
    const ModConst m = S.mc[L.idx[row]];
    const size_t er = ERow.idx[row];
    const size_t N = (size_t)1 << S.logn;
    const int jown = row < nq ? row / alpha : -1;
    FOR_THREADS {
        const u32 k = blockIdx.x * TPB + threadIdx.x;
        const u32 ks = perm ? ldg(perm + k) : k;
        // the key words of this coefficient: loaded once, applied to every batch item
        u64 e0[KS_MAX_BETA], e1[KS_MAX_BETA];
#pragma unroll
        for (int j = 0; j < KS_MAX_BETA; j++) {
            e0[j] = e1[j] = 0;
            if (j < beta) {
                const u64* e = evk + ((size_t)j * 2 * evk_rows + er) * N + k;
                e0[j] = ldg(e);
                e1[j] = ldg(e + (size_t)evk_rows * N);
            }
        }
        for (int bi = 0; bi < kb.nb; bi++) {
            const u64* xext = ext + bi * kb.ext;
            const u64* xown = own + bi * kb.own;
            const u64* xadd = addend ? addend + bi * kb.addend : nullptr;
            u64* xacc = acc + bi * kb.acc;
            u64 h0 = 0, l0 = 0, h1 = 0, l1 = 0;
            u64 a0 = 0, a1 = 0, b0 = 0, b1 = 0;
            if (TENSOR && row < nq) {
                a0 = xown[(size_t)row * N + k];  a1 = xown[((size_t)nq + row) * N + k];
                b0 = xadd[(size_t)row * N + k];  b1 = xadd[((size_t)nq + row) * N + k];
            }
#pragma unroll
            for (int j = 0; j < KS_MAX_BETA; j++) {
                if (j < beta) {
                    const u64 x = j == jown ? (TENSOR ? barrett_mul(a1, b1, m) : xown[(size_t)row * N + ks])
                                            : xext[((size_t)j * rows + row) * N + ks];
                    mac128(h0, l0, x, e0[j]);
                    mac128(h1, l1, x, e1[j]);
                    if ((j & 3) == 3 && j + 1 < beta) {      // keep the 128-bit sums below 4 q^2 (61-bit moduli)
                        l0 = barrett_reduce128(h0, l0, m); h0 = 0;
                        l1 = barrett_reduce128(h1, l1, m); h1 = 0;
                    }
                }
            }
            u64 r0 = barrett_reduce128(h0, l0, m), r1 = barrett_reduce128(h1, l1, m);
            if (TENSOR) {
                if (row < nq) {
                    u64 hi = 0, lo = 0;
                    mac128(hi, lo, a0, b1);
                    mac128(hi, lo, a1, b0);
                    r0 = add_mod(r0, shoup_mul(barrett_mul(a0, b0, m), PmodQ.v[row], PmodQ.vs[row], m.q), m.q);
                    r1 = add_mod(r1, shoup_mul(barrett_reduce128(hi, lo, m), PmodQ.v[row], PmodQ.vs[row], m.q), m.q);
                }
            } else if (xadd != nullptr && row < nq) {
                r0 = add_mod(r0, shoup_mul(xadd[(size_t)row * N + (kb.addend_mode == 2 ? ks : k)], PmodQ.v[row], PmodQ.vs[row], m.q), m.q);
                if (kb.addend_mode == 0)
                    r1 = add_mod(r1, shoup_mul(xadd[((size_t)nq + row) * N + k], PmodQ.v[row], PmodQ.vs[row], m.q), m.q);
            }
            if (accumulate) {          // several key switches summed before ONE ModDown (giant steps of a linear transform)
                r0 = add_mod(r0, xacc[(size_t)row * N + k], m.q);
                r1 = add_mod(r1, xacc[((size_t)rows + row) * N + k], m.q);
            }
            xacc[(size_t)row * N + k] = r0;
            xacc[((size_t)rows + row) * N + k] = r1;
        }
    }
}

// The same inner product with the digit count and the number of batch items handled together as template parameters:
// every digit word of NB items (and the addend / accumulator words) is requested before the first multiply-accumulate,
// so a thread keeps up to NB * BETA + 2 BETA + 4 NB loads in flight instead of BETA (the run-time-sized loop above is
// bound by the latency of its loads: ncu long_scoreboard 7.9 per issue, 2.1 TB/s).  Same arithmetic, same results.
template <bool TENSOR, int NB, int BETA>
__global__ void __launch_bounds__(TPB)
k_ks_inner_mlp(KShape S, u64* __restrict__ acc, const u64* __restrict__ ext, const u64* __restrict__ own,
               const u64* __restrict__ evk, const u32* __restrict__ perm, const GRID_CONST LimbList L,
               const GRID_CONST LimbList ERow, int rows, int evk_rows, int nq, int alpha,
               const u64* __restrict__ addend, const GRID_CONST ScalarList PmodQ, int accumulate,
               const GRID_CONST KsBatch kb) {
    static_assert(BETA >= 1 && BETA <= 4, "no intermediate reduction of the 128-bit sums below five digits");
    const int row = blockIdx.y;
    const ModConst m = S.mc[L.idx[row]];
    const size_t er = ERow.idx[row];
    const size_t N = (size_t)1 << S.logn;
    const int jown = row < nq ? row / alpha : -1;
    const bool qrow = row < nq;
#ifndef KS_INNER_FP
#define KS_INNER_FP 1         // 0: 128-bit integer multiply-accumulates on every row (A/B)
#endif
    const bool fprow = KS_INNER_FP && m.q < CKKS_FP_LIMIT;          // uniform over the CTA (one row per blockIdx.y)
    FOR_THREADS {
        const u32 k = blockIdx.x * TPB + threadIdx.x;
        const u32 ks = perm ? ldg(perm + k) : k;
        u64 e0[BETA], e1[BETA];
#pragma unroll
        for (int j = 0; j < BETA; j++) {
            const u64* e = evk + ((size_t)j * 2 * evk_rows + er) * N + k;
            e0[j] = ldg(e);
            e1[j] = ldg(e + (size_t)evk_rows * N);
        }
        double ed0[BETA], ed1[BETA];
        const double qd = ull2d_rn(m.q), qinv = fdiv_rn(1.0, qd);
        if (fprow) {
#pragma unroll
            for (int j = 0; j < BETA; j++) { ed0[j] = ull2d_rn(e0[j]); ed1[j] = ull2d_rn(e1[j]); }
        }
        for (int b0 = 0; b0 < kb.nb; b0 += NB) {
            u64 x[NB][BETA], p0[NB], p1[NB], p2[NB], p3[NB], c0[NB], c1[NB];
#pragma unroll
            for (int bi = 0; bi < NB; bi++) {
                const int b = b0 + bi < kb.nb ? b0 + bi : kb.nb - 1;        // a short last group repeats its last item
                const u64* xext = ext + b * kb.ext;
                const u64* xown = own + b * kb.own;
                const u64* xadd = addend ? addend + b * kb.addend : nullptr;
                const u64* xacc = acc + b * kb.acc;
                p0[bi] = p1[bi] = p2[bi] = p3[bi] = c0[bi] = c1[bi] = 0;
#pragma unroll
                for (int j = 0; j < BETA; j++) {
                    if (j == jown) x[bi][j] = TENSOR ? 0 : xown[(size_t)row * N + ks];
                    else x[bi][j] = xext[((size_t)j * rows + row) * N + ks];
                }
                if (TENSOR) {
                    if (qrow) {
                        p0[bi] = xown[(size_t)row * N + k];  p1[bi] = xown[((size_t)nq + row) * N + k];
                        p2[bi] = xadd[(size_t)row * N + k];  p3[bi] = xadd[((size_t)nq + row) * N + k];
                    }
                } else if (xadd != nullptr && qrow) {
                    p0[bi] = xadd[(size_t)row * N + (kb.addend_mode == 2 ? ks : k)];
                    p1[bi] = kb.addend_mode == 0 ? xadd[((size_t)nq + row) * N + k] : 0;
                }
                if (accumulate) {
                    c0[bi] = xacc[(size_t)row * N + k];
                    c1[bi] = xacc[((size_t)rows + row) * N + k];
                }
            }
#pragma unroll
            for (int bi = 0; bi < NB; bi++) {
                if (b0 + bi >= kb.nb) break;
                u64* xacc = acc + (b0 + bi) * kb.acc;
                u64 r0, r1;
                if (fprow) {
                    // FP64 pipe: the digit words and the key words are below q < 2^50.2, every product an exact FP64
                    // modular product (common.cuh), the sum of BETA <= 4 of them far below 2^53
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int j = 0; j < BETA; j++) {
                        const double xd = ull2d_rn((TENSOR && j == jown) ? barrett_mul(p1[bi], p3[bi], m) : x[bi][j]);
                        s0 = fadd_rn(s0, modmul_fp(xd, ed0[j], qd, qinv));
                        s1 = fadd_rn(s1, modmul_fp(xd, ed1[j], qd, qinv));
                    }
                    r0 = canon_fp(s0, qd, qinv);
                    r1 = canon_fp(s1, qd, qinv);
                } else {
                    u64 h0 = 0, l0 = 0, h1 = 0, l1 = 0;
#pragma unroll
                    for (int j = 0; j < BETA; j++) {
                        const u64 xv = (TENSOR && j == jown) ? barrett_mul(p1[bi], p3[bi], m) : x[bi][j];
                        mac128(h0, l0, xv, e0[j]);
                        mac128(h1, l1, xv, e1[j]);
                    }
                    r0 = barrett_reduce128(h0, l0, m);
                    r1 = barrett_reduce128(h1, l1, m);
                }
                if (TENSOR) {
                    if (qrow) {
                        u64 hi = 0, lo = 0;
                        mac128(hi, lo, p0[bi], p3[bi]);
                        mac128(hi, lo, p1[bi], p2[bi]);
                        r0 = add_mod(r0, shoup_mul(barrett_mul(p0[bi], p2[bi], m), PmodQ.v[row], PmodQ.vs[row], m.q), m.q);
                        r1 = add_mod(r1, shoup_mul(barrett_reduce128(hi, lo, m), PmodQ.v[row], PmodQ.vs[row], m.q), m.q);
                    }
                } else if (addend != nullptr && qrow) {
                    r0 = add_mod(r0, shoup_mul(p0[bi], PmodQ.v[row], PmodQ.vs[row], m.q), m.q);
                    if (kb.addend_mode == 0) r1 = add_mod(r1, shoup_mul(p1[bi], PmodQ.v[row], PmodQ.vs[row], m.q), m.q);
                }
                if (accumulate) {
                    r0 = add_mod(r0, c0[bi], m.q);
                    r1 = add_mod(r1, c1[bi], m.q);
                }
                xacc[(size_t)row * N + k] = r0;
                xacc[((size_t)rows + row) * N + k] = r1;
            }
        }
    }
}

// ------------------------------------------------------------------ fast basis conversion (spec S5)
// in: coefficient-domain residues, source i in row T.srow[i]; out row T.orow[t] for target t.
// One thread per coefficient computes ALL targets: y_i = x_i * hatinv_i mod s_i once, then for every target a
// 128-bit multiply-accumulate sum_i y_i * hat[i][t] with ONE reduction at the end (no per-term reduction, no Shoup
// companion loads).  The hat matrix and 2^64 mod q_t sit in shared memory (broadcast reads).  NS is a template
// parameter so the source loop is fully unrolled without predicated-off iterations.  Integer-pipe bound.
template <int NS>
__global__ void __launch_bounds__(TPB)
k_base_convert(KShape S, u64* __restrict__ out, const u64* __restrict__ in, const BaseConvTable* __restrict__ tabs,
               int tab_zstride, size_t in_zs, size_t out_zs, int nz, size_t in_bs, size_t out_bs) {
    CKKS_SHARED u64 s_hat[BC_MAX_TGT * NS];
    CKKS_SHARED u64 s_r64[BC_MAX_TGT];
    const size_t N = (size_t)1 << S.logn;
    const unsigned zb = blockIdx.z / (unsigned)nz, zz = blockIdx.z - zb * (unsigned)nz;     // batch item, slice
    const BaseConvTable& T = tabs[zz * tab_zstride];              // slice z uses its own table (digit) or a shared one
    const int nt = T.nt;
    // blockIdx.y picks a chunk of BC_CHUNK targets: more CTAs in flight for the small launches of a key switch; the
    // NS multiplications for y are repeated per chunk (NS of NS*(BC_CHUNK+1))
    const int tbeg = blockIdx.y * BC_CHUNK, tend = tbeg + BC_CHUNK < nt ? tbeg + BC_CHUNK : nt;
    if (tbeg >= nt) return;
    FOR_THREADS {
        for (int e = threadIdx.x; e < nt * NS; e += TPB) {
            const int t = e / NS, i = e - t * NS;
            s_hat[e] = ldg(T.hat + i * nt + t);
        }
        for (int t = threadIdx.x; t < nt; t += TPB) {
            const ModConst m = S.mc[T.tgt[t]];
            s_r64[t] = barrett_reduce128(1, 0, m);                // 2^64 mod q_t
        }
    }
    BLOCK_SYNC;
    FOR_THREADS {
        const u32 k = blockIdx.x * TPB + threadIdx.x;
        const u64* src = in + zz * in_zs + zb * in_bs;
        u64* dst = out + zz * out_zs + zb * out_bs;
        u64 y[NS];
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < NS; i++) {
            y[i] = shoup_mul(src[(size_t)T.srow[i] * N + k], T.hatinv[i], T.hatinv_s[i], S.mc[T.src[i]].q);
            v = fadd_rn(v, fmul_rn(ull2d_rn(y[i]), T.inv_src[i]));
        }
        const u64 u = T.exact ? (u64)d2ll_rn(v) : 0;       // overflow count of the fast conversion (0..NS)
        for (int t = tbeg; t < tend; t++) {
            const ModConst m = S.mc[T.tgt[t]];
            u64 hi = 0, lo = 0;
#pragma unroll
            for (int i = 0; i < NS; i++) mac128(hi, lo, y[i], s_hat[t * NS + i]);
            if (T.exact) mac128(hi, lo, u, T.negD[t]);      // - u * D  (mod q_t)
            // (hi, lo) < NS * 2^122: fold the high word through 2^64 mod q_t, then one Barrett reduction
            const u64 h1 = barrett_reduce64(hi, m);
            u64 h2 = 0, l2 = lo;
            const u64 pl = h1 * s_r64[t];
            l2 += pl;
            h2 = mulhi64(h1, s_r64[t]) + (l2 < pl);
            dst[(size_t)T.orow[t] * N + k] = barrett_reduce128(h2, l2, m);
        }
    }
}

// ------------------------------------------------------------------ basis conversion on the FP64 pipe (+ the integer pipe)
// The same sums as k_base_convert, evaluated per target on the pipe that suits its modulus.  A target below
// CKKS_FP_LIMIT (the ~2^50 scale primes, and the special primes of the default chain) is a sum of exact FP64 modular
// products (common.cuh: modmul_fp, 5 FP64-pipe instructions + 1 rounding each, measured 11.1 per clock and SM against 4.8
// 128-bit integer multiply-accumulates):
//     out_t = canon( sum_i y_i * hat[i][t] mod q_t ),   residues of the products kept as exact integers in doubles.
// A source modulus of 60/61 bits (q_0; 61-bit special primes if a chain has them) gives y_i that a double cannot hold: it
// enters as y_i = yh 2^32 + yl with a second constant hat 2^32 mod q_t.  Two accumulators (even / odd sources) are folded
// every four sources: |acc| <= 0.51 q + 4 * 0.85 q stays far below 2^53.  Targets of 60/61 bits keep the 128-bit integer
// multiply-accumulate of k_base_convert; warps working on them run beside FP64 warps on the same SM.  Results are
// canonical residues, identical to k_base_convert's bit for bit (the arithmetic is exact on both paths).
// GENERIC = 0: only source 0 may be wide (q_0 in the first digit; warp-uniform, decided once per CTA): the source loop
// is branch-free.  GENERIC = 1: any source may be wide (run-time mask; the compiler predicates both forms).
#ifndef BC_FP_MIN_BLOCKS
#define BC_FP_MIN_BLOCKS 4
#endif
struct BcFpTarget { double q, qinv, negD, pad; };
// sum_i y_i * hat_i mod q_t as a lazy double: products in groups of five (tree of exact additions), one fold between
// groups.  Bounds in units of q_t <= 1.1 * 2^50 (2^53 = 7.27 q): a product of a narrow source is at most 0.92, of a half
// of a wide source 0.5; first group (+ the high half of source 0) <= 5.1, later: 0.51 + 4.6 = 5.11; the caller may add
// one more product of a small operand (0.5).
template <int NS, bool W0>
__device__ __forceinline__ double bc_fp_sum(const double (&yl)[NS], double yh0, const double* hf, const double* hw, double q,
                                            double qinv) {
    double s = 0.0;
#pragma unroll
    for (int g0 = 0; g0 < NS; g0 += 5) {
        double r[5];
#pragma unroll
        for (int j = 0; j < 5; j++)
            r[j] = g0 + j < NS ? modmul_fp(yl[g0 + j], hf[g0 + j], q, qinv) : 0.0;
        double p = fadd_rn(r[0], r[1]);
        if (g0 + 2 < NS) p = fadd_rn(p, g0 + 3 < NS ? fadd_rn(r[2], r[3]) : r[2]);
        if (g0 + 4 < NS) p = fadd_rn(p, r[4]);
        if (g0 == 0) s = W0 ? fadd_rn(p, modmul_fp(yh0, hw[0], q, qinv)) : p;
        else s = fadd_rn(fold_fp(s, q, qinv), p);
    }
    return s;
}
template <int NS, int GENERIC>
__global__ void __launch_bounds__(TPB, BC_FP_MIN_BLOCKS)
k_base_convert_fp(KShape S, u64* __restrict__ out, const u64* __restrict__ in, const BaseConvTable* __restrict__ tabs,
                  int tab_zstride, size_t in_zs, size_t out_zs, int nz, size_t in_bs, size_t out_bs) {
    constexpr int HW = GENERIC ? 2 : 1;                 // doubles per (target, source): hat (and hat 2^32 for wide sources)
    CKKS_SHARED u64 s_hat[BC_CHUNK * NS];
    CKKS_SHARED u64 s_r64[BC_CHUNK];
    CKKS_SHARED __align__(16) double s_hf[BC_CHUNK * NS * HW];
    CKKS_SHARED double s_hw[BC_CHUNK];                               // source 0's second constant (GENERIC = 0)
    CKKS_SHARED BcFpTarget s_tg[BC_CHUNK];
    CKKS_SHARED int s_nint;                                          // integer-pipe targets in this chunk
    const size_t N = (size_t)1 << S.logn;
    const unsigned zb = blockIdx.z / (unsigned)nz, zz = blockIdx.z - zb * (unsigned)nz;     // batch item, slice
    const BaseConvTable& T = tabs[zz * tab_zstride];
    const int nt = T.nt;
    const int tbeg = blockIdx.y * BC_CHUNK, tend = tbeg + BC_CHUNK < nt ? tbeg + BC_CHUNK : nt;
    if (tbeg >= nt) return;
    FOR_THREADS {
        const int cnt = (tend - tbeg) * NS;
        for (int e = threadIdx.x; e < cnt; e += TPB) {
            const int t = tbeg + e / NS, i = e % NS;
            s_hat[e] = ldg(T.hat + i * nt + t);
            const double* h = T.hatf + ((size_t)t * NS + i) * 2;
#pragma unroll
            for (int c = 0; c < HW; c++) s_hf[e * HW + c] = ldg(h + c);
            if (!GENERIC && i == 0) s_hw[t - tbeg] = ldg(h + 1);
        }
        for (int t = tbeg + threadIdx.x; t < tend; t += TPB) {
            const ModConst m = S.mc[T.tgt[t]];
            s_r64[t - tbeg] = barrett_reduce128(1, 0, m);         // 2^64 mod q_t
            BcFpTarget g;
            g.q = ull2d_rn(m.q); g.qinv = T.tqinv[t]; g.negD = T.negDd[t]; g.pad = 0.0;
            s_tg[t - tbeg] = g;
        }
        if (threadIdx.x == 0) {
            int c = 0;
            for (int t = tbeg; t < tend; t++) c += T.tfp[t] ? 0 : 1;
            s_nint = c;
        }
    }
    BLOCK_SYNC;
    FOR_THREADS {
        const u32 k = blockIdx.x * TPB + threadIdx.x;
        const u64* src = in + zz * in_zs + zb * in_bs;
        u64* dst = out + zz * out_zs + zb * out_bs;
        const u32 wide = T.swide;
        const int exact = T.exact;
        // the FP64 targets read the y_i as doubles only (low / high halves for a wide source); the integer residues are
        // rebuilt from them for the (few) integer-pipe targets, so the main loop does not carry both forms in registers
        double yl[NS], yh[GENERIC ? NS : 1];
        double v = 0.0;
        yh[0] = 0.0;
#pragma unroll
        for (int i = 0; i < NS; i++) {
            const u64 y = shoup_mul(src[(size_t)T.srow[i] * N + k], T.hatinv[i], T.hatinv_s[i], S.mc[T.src[i]].q);
            const double yd = ull2d_rn(y);
            v = fadd_rn(v, fmul_rn(yd, T.inv_src[i]));
            if ((GENERIC || i == 0) && (wide >> i & 1)) {
                yl[i] = ull2d_rn(y & 0xffffffffull);
                yh[GENERIC ? i : 0] = ull2d_rn(y >> 32);
            } else {
                yl[i] = yd;
                if (GENERIC) yh[i] = 0.0;
            }
        }
        const double ud = exact ? frint(v) : 0.0;        // overflow count of the fast conversion (0..NS)
        for (int t = tbeg; t < tend; t++) {
            if (!T.tfp[t]) continue;
            const BcFpTarget g = s_tg[t - tbeg];
            const double* hf = s_hf + (t - tbeg) * NS * HW;
            double s;
            if (GENERIC) {
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int i = 0; i < NS; i++) {
                    if (wide >> i & 1) {
                        a0 = fadd_rn(a0, modmul_fp(yl[i], hf[2 * i], g.q, g.qinv));
                        a1 = fadd_rn(a1, modmul_fp(yh[GENERIC ? i : 0], hf[2 * i + 1], g.q, g.qinv));
                    } else if (i & 1) {
                        a1 = fadd_rn(a1, modmul_fp(yl[i], hf[2 * i], g.q, g.qinv));
                    } else {
                        a0 = fadd_rn(a0, modmul_fp(yl[i], hf[2 * i], g.q, g.qinv));
                    }
                    if ((i & 3) == 3 && i + 1 < NS) { a0 = fold_fp(a0, g.q, g.qinv); a1 = fold_fp(a1, g.q, g.qinv); }
                }
                s = fadd_rn(fold_fp(a0, g.q, g.qinv), fold_fp(a1, g.q, g.qinv));
            } else if (wide & 1) {
                s = bc_fp_sum<NS, true>(yl, yh[0], hf, s_hw + (t - tbeg), g.q, g.qinv);
            } else {
                s = bc_fp_sum<NS, false>(yl, 0.0, hf, nullptr, g.q, g.qinv);
            }
            if (exact) s = fadd_rn(s, modmul_fp(ud, g.negD, g.q, g.qinv));             // - u * D  (mod q_t)
            dst[(size_t)T.orow[t] * N + k] = canon_fp(s, g.q, g.qinv);
        }
        if (s_nint) {
            u64 y[NS];
#pragma unroll
            for (int i = 0; i < NS; i++) {
                y[i] = (u64)d2ll_rn(yl[i]);
                if ((GENERIC || i == 0) && (wide >> i & 1)) y[i] |= (u64)d2ll_rn(yh[GENERIC ? i : 0]) << 32;
            }
            const u64 u = (u64)d2ll_rn(ud);
            for (int t = tbeg; t < tend; t++) {
                if (T.tfp[t]) continue;
                const ModConst m = S.mc[T.tgt[t]];
                const u64* hat = s_hat + (t - tbeg) * NS;
                u64 hi = 0, lo = 0;
#pragma unroll
                for (int i = 0; i < NS; i++) mac128(hi, lo, y[i], hat[i]);
                if (exact) mac128(hi, lo, u, T.negD[t]);
                const u64 r64 = s_r64[t - tbeg];
                const u64 h1 = barrett_reduce64(hi, m);
                u64 h2 = 0, l2 = lo;
                const u64 pl = h1 * r64;
                l2 += pl;
                h2 = mulhi64(h1, r64) + (l2 < pl);
                dst[(size_t)T.orow[t] * N + k] = barrett_reduce128(h2, l2, m);
            }
        }
    }
}

#ifndef CKKS_EMU
// ------------------------------------------------------------------ basis conversion on the tensor cores
// The same sum  out[n][t] = sum_i y_i[n] * hat[i][t]  (mod q_t), as an 8-bit integer GEMM.  y_i and hat are cut into
// bytes, y_i = sum_a ya 2^(8a), hat = sum_b hb 2^(8b); the products with a + b = d form diagonal d:
//     S_d[n][t] = sum_{i, a} ya[n][(i,a)] * hb[(i,a)][(t,d)],   hb[(i,a)][(t,d)] = byte (d - a) of hat[i][t]  (or 0)
// i.e. ONE [16 d x 8 t] x [64] x [n] integer GEMM per group of 8 targets (mma.sync.m16n8k32.u8.u8.s32; |S_d| < 2^22).
// The M rows of tile j are ordered (d = 2j, t = 0..7), (d = 2j + 1, t = 0..7), so the thread with groupID g ends up holding
// ALL 16 diagonals of target g for its two coefficients and finishes them alone:
//     out = sum_d S_d * (2^(8d) mod q_t)  (< 2^88: two 32 x 32 + 64 multiply-adds per diagonal)  -> ONE Barrett reduction.
// About 70 CUDA-core instructions per output instead of ~195 (8 x 128-bit multiply-accumulate + three-step reduction).
__device__ __forceinline__ void mma_u8(int (&c)[4], const uint4& a, u32 b0, u32 b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
constexpr int BCM_TILE = 128;          // coefficients per CTA: 4 warps x 4 steps x 8
constexpr int BCM_YSTRIDE = 9;         // u64 per coefficient row of ys (8 sources + 1 pad: conflict-free 32-bit reads)
#ifndef BCM_MIN_BLOCKS
#define BCM_MIN_BLOCKS 6
#endif
#ifndef BCM_TG_PER_CTA
#define BCM_TG_PER_CTA 8      // all target groups in one CTA (1 or 2 per CTA measured slower: y_i recomputed, more CTAs than SM slots)
#endif
template <int NS>
__global__ void __launch_bounds__(BCM_TILE, BCM_MIN_BLOCKS)
k_base_convert_mma(KShape S, u64* __restrict__ out, const u64* __restrict__ in, const BaseConvTable* __restrict__ tabs,
                   int tab_zstride, size_t in_zs, size_t out_zs, int nz, size_t in_bs, size_t out_bs) {
    __shared__ __align__(16) u64 ys[BCM_TILE * BCM_YSTRIDE];
    __shared__ u64 s_pow8[BC_MAX_TGT * 16];
    __shared__ u64 s_q[BC_MAX_TGT], s_mu[BC_MAX_TGT], s_negD[BC_MAX_TGT];
    __shared__ u32 s_k1[BC_MAX_TGT], s_orow[BC_MAX_TGT];
    __shared__ unsigned char s_u[BCM_TILE];
    const size_t N = (size_t)1 << S.logn;
    const unsigned zb = blockIdx.z / (unsigned)nz, zz = blockIdx.z - zb * (unsigned)nz;
    const BaseConvTable& T = tabs[zz * tab_zstride];
    const int nt = T.nt, tid = threadIdx.x;
    // blockIdx.y picks BCM_TG_PER_CTA target groups: more CTAs in flight for the small launches of a key switch (the y_i
    // are recomputed per CTA: NS multiplications against 8 x 16 diagonals of recombination)
    const int tg_beg = blockIdx.y * BCM_TG_PER_CTA;
    if (tg_beg * 8 >= nt) return;
    const u64* src = in + zz * in_zs + zb * in_bs;
    u64* dst = out + zz * out_zs + zb * out_bs + (size_t)blockIdx.x * BCM_TILE;
    {
        // y_i = x_i * (D/s_i)^-1 mod s_i and the overflow count of the exact variant (same arithmetic as k_base_convert)
        const size_t k = (size_t)blockIdx.x * BCM_TILE + tid;
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            u64 y = 0;
            if (i < NS) {
                y = shoup_mul(src[(size_t)T.srow[i] * N + k], T.hatinv[i], T.hatinv_s[i], S.mc[T.src[i]].q);
                v = fadd_rn(v, fmul_rn(ull2d_rn(y), T.inv_src[i]));
            }
            ys[tid * BCM_YSTRIDE + i] = y;
        }
        s_u[tid] = T.exact ? (unsigned char)d2ll_rn(v) : 0;
        for (int e = tid; e < nt * 16; e += BCM_TILE) s_pow8[e] = ldg(T.pow8 + e);
        if (tid < nt) {
            const ModConst m = S.mc[T.tgt[tid]];
            s_q[tid] = m.q; s_mu[tid] = m.mu; s_k1[tid] = m.k1;
            s_negD[tid] = T.negD[tid];
            s_orow[tid] = T.orow[tid];
        }
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, q4 = lane & 3;
    const uint4* AF = reinterpret_cast<const uint4*>(T.afrag) + lane;
    const u32* yw = reinterpret_cast<const u32*>(ys);
    const int exact = T.exact;
#pragma unroll 1
    for (int step = 0; step < 4; step++) {
        const int n0 = warp * 32 + step * 8;
        // B fragment: bytes k = (i, a) of coefficient n0 + g; this lane owns the 32-bit half q4 & 1 of sources
        // q4/2, 2 + q4/2 (k-step 0) and 4 + q4/2, 6 + q4/2 (k-step 1)
        const u32* row = yw + (size_t)(n0 + g) * (2 * BCM_YSTRIDE) + (q4 & 1);
        const u32 b00 = row[2 * (q4 >> 1)], b01 = row[2 * (2 + (q4 >> 1))];
        const u32 b10 = row[2 * (4 + (q4 >> 1))], b11 = row[2 * (6 + (q4 >> 1))];
        const u64 u0 = s_u[n0 + 2 * q4], u1 = s_u[n0 + 2 * q4 + 1];
#pragma unroll 1
        for (int tg = tg_beg; tg < tg_beg + BCM_TG_PER_CTA && tg * 8 < nt; tg++) {
            // the hat fragments of this target group come straight from the (L1-resident, 8 KB per group) table: holding
            // them in registers across the steps costs 64 registers and half the resident CTAs
            int acc[8][4];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0;
                mma_u8(acc[j], __ldg(AF + ((size_t)(tg * 8 + j) * 2) * 32), b00, b01);
                if (NS > 4) mma_u8(acc[j], __ldg(AF + ((size_t)(tg * 8 + j) * 2 + 1) * 32), b10, b11);
            }
            const int t = tg * 8 + g;
            if (t < nt) {
                ModConst m;
                m.q = s_q[t]; m.mu = s_mu[t]; m.k1 = s_k1[t];
                const u64* pw = s_pow8 + t * 16;
                u64 a00 = 0, a01 = 0, a10 = 0, a11 = 0;
#pragma unroll
                for (int d = 0; d < 15; d++) {
                    const u64 p = pw[d], pl = p & 0xffffffffull, ph = p >> 32;
                    const u64 s0 = (u64)(u32)acc[d >> 1][(d & 1) * 2], s1 = (u64)(u32)acc[d >> 1][(d & 1) * 2 + 1];
                    a00 += s0 * pl; a01 += s0 * ph;
                    a10 += s1 * pl; a11 += s1 * ph;
                }
                u64 lo0 = a00 + (a01 << 32), hi0 = (a01 >> 32) + (lo0 < a00);
                u64 lo1 = a10 + (a11 << 32), hi1 = (a11 >> 32) + (lo1 < a10);
                if (exact) { mac128(hi0, lo0, u0, s_negD[t]); mac128(hi1, lo1, u1, s_negD[t]); }
                *reinterpret_cast<ulonglong2*>(dst + (size_t)s_orow[t] * N + n0 + 2 * q4) =
                    make_ulonglong2(barrett_reduce128(hi0, lo0, m), barrett_reduce128(hi1, lo1, m));
            }
        }
    }
}
#endif

// ------------------------------------------------------------------ rescale helpers (spec S6)
// t = (last + h) mod q_l (coefficient domain, single limb) -> delta[i] = (t mod q_i) - (h mod q_i)
__global__ void k_rescale_delta(KShape S, u64* __restrict__ delta, const u64* __restrict__ last, const GRID_CONST LimbList L,
                                int last_mod, PolyStride ps) {
    const int row = blockIdx.y;
    const ModConst m = S.mc[L.idx[row]];
    const u64 ql = S.mc[last_mod].q;
    const u64 h = ql >> 1;
    const ZOff o = zoff(ps);
    FOR_THREADS {
        const u32 k = blockIdx.x * TPB + threadIdx.x;
        u64 t = last[k + o.a] + h;
        t = t >= ql ? t - ql : t;
        const u64 r = barrett_reduce64(t, m), hm = barrett_reduce64(h, m);
        delta[((size_t)row << S.logn) + k + o.out] = sub_mod(r, hm, m.q);
    }
}

// centred lift of a single coefficient-domain limb (mod q_src) into every row modulus (ModRaise)
__global__ void k_center_lift(KShape S, u64* __restrict__ out, const u64* __restrict__ in, const GRID_CONST LimbList L, int src_mod,
                              PolyStride ps) {
    const int row = blockIdx.y;
    const ModConst m = S.mc[L.idx[row]];
    const u64 qs = S.mc[src_mod].q;
    const ZOff o = zoff(ps);
    FOR_THREADS {
        const u32 k = blockIdx.x * TPB + threadIdx.x;
        const u64 v = in[k + o.a];
        u64 r;
        if (v > (qs >> 1)) r = neg_mod(barrett_reduce64(qs - v, m), m.q);
        else r = barrett_reduce64(v, m);
        out[((size_t)row << S.logn) + k + o.out] = r;
    }
}

// ------------------------------------------------------------------ sampling (spec S8)
__device__ __forceinline__ u64 mix64(u64 z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ u64 rand64(u64 seed, u64 stream, u64 idx) {
    return mix64(mix64(seed + stream * 0xD1342543DE82EF95ull) + idx);
}
__global__ void k_sample_uniform(KShape S, u64* __restrict__ out, const GRID_CONST LimbList L, u64 seed, u64 stream) {
    const int row = blockIdx.y;
    const int mod = L.idx[row];
    const u64 q = S.mc[mod].q;
    const u64 bound = (u64)0 - ((u64)0 - q) % q;
    FOR_THREADS {
        const u64 k = blockIdx.x * TPB + threadIdx.x;
        u64 r = 0;
        for (int t = 0; t < 8; t++) {
            r = rand64(seed, stream, ((((u64)mod) << S.logn) + k) * 8 + t);
            if (bound == 0 || r < bound) break;
        }
        out[((size_t)row << S.logn) + k] = r % q;
    }
}
// kind 0: centred binomial (21+21 bits); kind 1: ternary {-1,0,1} w.p. 1/4,1/2,1/4.  Same value in every row.
// blockIdx.z = batch item: its own stream (the "a" field of the stream id advances by one per item) and output slice.
// epoch != null: the seed is offset by the replay epoch of captured graphs (see launch_sample_small).
__global__ void k_sample_small(KShape S, u64* __restrict__ out, const GRID_CONST LimbList L, u64 seed, u64 stream, int kind,
                               size_t out_bs, const u64* __restrict__ epoch) {
    const int row = blockIdx.y;
    const u64 q = S.mc[L.idx[row]].q;
    const u64 sd = seed + (epoch ? *epoch : 0) * 0xA24BAED4963EE407ull;
    const u64 str = stream + ((u64)blockIdx.z << 16);
    out += blockIdx.z * out_bs;
    FOR_THREADS {
        const u64 k = blockIdx.x * TPB + threadIdx.x;
        const u64 r = rand64(sd, str, k);
        int v = kind == 0 ? popc64(r & 0x1FFFFF) - popc64((r >> 21) & 0x1FFFFF) : (int)(r & 1) - (int)((r >> 1) & 1);
        out[((size_t)row << S.logn) + k] = v < 0 ? q - (u64)(-v) : (u64)v;
    }
}
__global__ void k_bump(u64* __restrict__ word) {
    FOR_THREADS { if (threadIdx.x == 0) *word += 1; }
}
// signed 64-bit coefficients -> residues in every row
__global__ void k_reduce_i64(KShape S, u64* __restrict__ out, const i64* __restrict__ v, const GRID_CONST LimbList L, size_t out_bs) {
    const int row = blockIdx.y;
    const ModConst m = S.mc[L.idx[row]];
    out += blockIdx.z * out_bs;                       // blockIdx.z = batch item (N coefficients each in v)
    v += (size_t)blockIdx.z << S.logn;
    FOR_THREADS {
        const u32 k = blockIdx.x * TPB + threadIdx.x;
        const i64 x = v[k];
        const u64 r = barrett_reduce64(x < 0 ? (u64)(-x) : (u64)x, m);
        out[((size_t)row << S.logn) + k] = (x < 0) ? neg_mod(r, m.q) : r;
    }
}

// ------------------------------------------------------------------ canonical embedding (spec S9), fp64, no FMA contraction
// one butterfly stage of the "special FFT" on n = N/2 complex values (v: interleaved re,im)
__global__ void k_fft_stage(KShape S, double* __restrict__ v, const u32* __restrict__ rot,
                            const double* __restrict__ ksi, int len, int inverse) {
    v += (size_t)blockIdx.y << S.logn;                // blockIdx.y = batch item (n complex = N doubles each)
    FOR_THREADS {
        const int b = blockIdx.x * TPB + threadIdx.x;     // butterfly index in [0, n/2)
        const int lenh = len >> 1, lenq = len << 2, gap = (2 << S.logn) / lenq;
        const int j = b % lenh, i = (b / lenh) * len;
        const int p0 = i + j, p1 = p0 + lenh;
        const int rj = rot[j] % lenq;
        const int idx = (inverse ? (lenq - rj) : rj) * gap;
        const double wr = ksi[2 * idx], wi = ksi[2 * idx + 1];
        const double ar = v[2 * p0], ai = v[2 * p0 + 1], br = v[2 * p1], bi = v[2 * p1 + 1];
        if (!inverse) {
            const double tr = fsub_rn(fmul_rn(br, wr), fmul_rn(bi, wi));
            const double ti = fadd_rn(fmul_rn(br, wi), fmul_rn(bi, wr));
            v[2 * p0] = fadd_rn(ar, tr); v[2 * p0 + 1] = fadd_rn(ai, ti);
            v[2 * p1] = fsub_rn(ar, tr); v[2 * p1 + 1] = fsub_rn(ai, ti);
        } else {
            const double dr = fsub_rn(ar, br), di = fsub_rn(ai, bi);
            v[2 * p0] = fadd_rn(ar, br); v[2 * p0 + 1] = fadd_rn(ai, bi);
            v[2 * p1] = fsub_rn(fmul_rn(dr, wr), fmul_rn(di, wi));
            v[2 * p1 + 1] = fadd_rn(fmul_rn(dr, wi), fmul_rn(di, wr));
        }
    }
}
// The same transform in two launches instead of one per stage (butterflies of one stage are independent, so the values
// are bit-identical to the stage-by-stage kernels and to the oracle).  n = N/2 complex points are split as hi * FFT_LOW + lo:
//   k_fft_low : the stages len <= FFT_LOW stay inside one contiguous block of FFT_LOW points -> one CTA, shared memory;
//               forward: bit-reversal gather on load, stages len = 2 .. FFT_LOW; inverse: stages FFT_LOW .. 2, bit-reversal
//               scatter and the 1/n scaling on store;
//   k_fft_high: the stages len > FFT_LOW couple points FFT_LOW apart -> one thread per lo holds the n / FFT_LOW values in
//               registers (16 at N = 2^16; no such stage at N = 2^12).
constexpr int FFT_LOW = 2048;
constexpr int FFT_HI_MAX = 16;
__device__ __forceinline__ void fft_butterfly(double& ar, double& ai, double& br, double& bi, double wr, double wi, int inverse) {
    if (!inverse) {
        const double tr = fsub_rn(fmul_rn(br, wr), fmul_rn(bi, wi));
        const double ti = fadd_rn(fmul_rn(br, wi), fmul_rn(bi, wr));
        const double xr = ar, xi = ai;
        ar = fadd_rn(xr, tr); ai = fadd_rn(xi, ti);
        br = fsub_rn(xr, tr); bi = fsub_rn(xi, ti);
    } else {
        const double dr = fsub_rn(ar, br), di = fsub_rn(ai, bi);
        ar = fadd_rn(ar, br); ai = fadd_rn(ai, bi);
        br = fsub_rn(fmul_rn(dr, wr), fmul_rn(di, wi));
        bi = fadd_rn(fmul_rn(dr, wi), fmul_rn(di, wr));
    }
}
__global__ void __launch_bounds__(FFT_LOW / 2)
k_fft_low(KShape S, double* __restrict__ out, const double* __restrict__ in, const u32* __restrict__ rot,
          const double* __restrict__ ksi, int inverse, double mul) {
    CKKS_SHARED double sm[2 * FFT_LOW];
    const int n = 1 << (S.logn - 1), bits = S.logn - 1;
    const int blk = n < FFT_LOW ? n : FFT_LOW;                // points of this CTA (n >= 2048 for both built rings)
    const size_t vec = (size_t)blockIdx.y << S.logn;          // batch item
    const int base = blockIdx.x * blk;
    FOR_THREADS {
        for (int e = threadIdx.x; e < blk; e += blockDim.x) {
            const int p = base + e;
            const int src = inverse ? p : (int)(brev32((u32)p) >> (32 - bits));
            sm[2 * e] = in[vec + 2 * (size_t)src];
            sm[2 * e + 1] = in[vec + 2 * (size_t)src + 1];
        }
    }
    BLOCK_SYNC;
    for (int st = 0; (2 << st) <= blk; st++) {
        const int len = inverse ? (blk >> st) : (2 << st);
        const int lenh = len >> 1, lenq = len << 2, gap = (2 << S.logn) / lenq;
        FOR_THREADS {
            for (int b = threadIdx.x; b < blk / 2; b += blockDim.x) {
                const int j = b % lenh, p0 = (b / lenh) * len + j, p1 = p0 + lenh;
                const int rj = rot[j] % lenq;
                const int idx = (inverse ? (lenq - rj) : rj) * gap;
                fft_butterfly(sm[2 * p0], sm[2 * p0 + 1], sm[2 * p1], sm[2 * p1 + 1], ksi[2 * idx], ksi[2 * idx + 1], inverse);
            }
        }
        BLOCK_SYNC;
    }
    FOR_THREADS {
        for (int e = threadIdx.x; e < blk; e += blockDim.x) {
            const int p = base + e;
            const int dst = inverse ? (int)(brev32((u32)p) >> (32 - bits)) : p;
            out[vec + 2 * (size_t)dst] = inverse ? fmul_rn(sm[2 * e], mul) : sm[2 * e];
            out[vec + 2 * (size_t)dst + 1] = inverse ? fmul_rn(sm[2 * e + 1], mul) : sm[2 * e + 1];
        }
    }
}
// NH = n / FFT_LOW values per thread, INV: direction -- both compile-time so the register arrays are indexed statically
template <int NH, int INV>
__global__ void k_fft_high(KShape S, double* __restrict__ v, const u32* __restrict__ rot, const double* __restrict__ ksi) {
    v += (size_t)blockIdx.y << S.logn;
    FOR_THREADS {
        const int lo = blockIdx.x * TPB + threadIdx.x;        // lo < FFT_LOW
        double xr[NH], xi[NH];
#pragma unroll
        for (int h = 0; h < NH; h++) { xr[h] = v[2 * ((size_t)h * FFT_LOW + lo)]; xi[h] = v[2 * ((size_t)h * FFT_LOW + lo) + 1]; }
#pragma unroll
        for (int st = 0; (2 << st) <= NH; st++) {
            const int hl = INV ? (NH >> st) : (2 << st);      // stage length in units of FFT_LOW
            const int hh = hl >> 1;
            const int len = hl * FFT_LOW, lenq = len << 2, gap = (2 << S.logn) / lenq;
#pragma unroll
            for (int b = 0; b < NH / 2; b++) {
                const int jh = b % hh, h0 = (b / hh) * hl + jh, h1 = h0 + hh;
                const int j = jh * FFT_LOW + lo;
                const int rj = rot[j] % lenq;
                const int idx = (INV ? (lenq - rj) : rj) * gap;
                fft_butterfly(xr[h0], xi[h0], xr[h1], xi[h1], ksi[2 * idx], ksi[2 * idx + 1], INV);
            }
        }
#pragma unroll
        for (int h = 0; h < NH; h++) { v[2 * ((size_t)h * FFT_LOW + lo)] = xr[h]; v[2 * ((size_t)h * FFT_LOW + lo) + 1] = xi[h]; }
    }
}
__global__ void k_bitrev_copy(KShape S, double* __restrict__ out, const double* __restrict__ in, double mul) {
    out += (size_t)blockIdx.y << S.logn;
    in += (size_t)blockIdx.y << S.logn;
    FOR_THREADS {
        const u32 i = blockIdx.x * TPB + threadIdx.x;      // n = N/2 complex values
        const u32 j = brev32(i) >> (32 - (S.logn - 1));
        out[2 * j] = fmul_rn(in[2 * i], mul);
        out[2 * j + 1] = fmul_rn(in[2 * i + 1], mul);
    }
}
// w (n complex) * scale, round half-even -> N signed coefficients; flags overflow
__global__ void k_round_coeffs(KShape S, i64* __restrict__ out, const double* __restrict__ w, double scale,
                               int* __restrict__ flag) {
    const u32 n = 1u << (S.logn - 1);
    out += (size_t)blockIdx.y << S.logn;
    w += (size_t)blockIdx.y << S.logn;
    FOR_THREADS {
        const u32 k = blockIdx.x * TPB + threadIdx.x;      // k < n
        const double a = fmul_rn(w[2 * k], scale), b = fmul_rn(w[2 * k + 1], scale);
        if (!(fabs(a) < 4.0e18) || !(fabs(b) < 4.0e18)) { *flag = 1; out[k] = 0; out[k + n] = 0; }
        else { out[k] = d2ll_rn(a); out[k + n] = d2ll_rn(b); }
    }
}
// centred lift of limb 0 coefficients -> w = coef / scale
__global__ void k_center_to_w(KShape S, double* __restrict__ w, const u64* __restrict__ coef, int mod, double scale) {
    const u32 n = 1u << (S.logn - 1);
    const u64 q = S.mc[mod].q, half = q >> 1;
    w += (size_t)blockIdx.y << S.logn;
    coef += (size_t)blockIdx.y << S.logn;
    FOR_THREADS {
        const u32 k = blockIdx.x * TPB + threadIdx.x;      // k < n
        const u64 a = coef[k], b = coef[k + n];
        const double da = a > half ? -(double)(q - a) : (double)a;
        const double db = b > half ? -(double)(q - b) : (double)b;
        w[2 * k] = fdiv_rn(da, scale);
        w[2 * k + 1] = fdiv_rn(db, scale);
    }
}

// hard renorm on the device: every slot is replaced by the nearest zeta_16 codeword exp(-2 pi i k / 16) (angle only,
// reference utils.py:15-19); with stride > 1 only slots j = 0 mod stride carry data and the others are set to 1.0
// (reference state_encoder.py:23-27)
__global__ void k_snap_zeta16(KShape S, double* __restrict__ z, const double* __restrict__ table, int stride) {
    z += (size_t)blockIdx.y << S.logn;
    FOR_THREADS {
        const u32 j = blockIdx.x * TPB + threadIdx.x;
        if (stride > 1 && (j % (u32)stride) != 0) { z[2 * j] = 1.0; z[2 * j + 1] = 0.0; }
        else {
            const double ang = atan2(z[2 * j + 1], z[2 * j]);
            int k = (int)rint(-ang * (16.0 / (2.0 * 3.14159265358979323846)));
            k = ((k % 16) + 16) % 16;
            z[2 * j] = table[2 * k];
            z[2 * j + 1] = table[2 * k + 1];
        }
    }
}

// zeta_16 codec on the device (the host side of reference utils.py:9-19 / state_encoder.py): nibble k <-> exp(-2 pi i k/16)
__global__ void k_zeta16_from_nibbles(KShape S, double* __restrict__ z, const unsigned char* __restrict__ nib,
                                      const double* __restrict__ table) {
    z += (size_t)blockIdx.y << S.logn;
    nib += (size_t)blockIdx.y << (S.logn - 1);
    FOR_THREADS {
        const u32 j = blockIdx.x * TPB + threadIdx.x;
        const int k = nib[j] & 15;
        z[2 * j] = table[2 * k];
        z[2 * j + 1] = table[2 * k + 1];
    }
}
__global__ void k_nibbles_from_zeta16(KShape S, unsigned char* __restrict__ nib, const double* __restrict__ z) {
    z += (size_t)blockIdx.y << S.logn;
    nib += (size_t)blockIdx.y << (S.logn - 1);
    FOR_THREADS {
        const u32 j = blockIdx.x * TPB + threadIdx.x;
        const double ang = atan2(z[2 * j + 1], z[2 * j]);
        int k = (int)rint(-ang * (16.0 / (2.0 * 3.14159265358979323846)));
        nib[j] = (unsigned char)(((k % 16) + 16) % 16);
    }
}

inline dim3 grid3(KShape S, int rows, int npoly = 1) { return dim3((1u << S.logn) / TPB, rows, npoly); }
// grid z = batch items x polynomials (batch-major); the kernels split blockIdx.z with ps.npoly
inline dim3 grid3b(KShape S, int rows, int npoly, PolyStride& ps) {
    ps.npoly = ps.nb > 1 ? npoly : 0;
    return dim3((1u << S.logn) / TPB, rows, npoly * (ps.nb > 1 ? ps.nb : 1));
}

}  // namespace

// ============================================================================ host wrappers
void launch_add(KShape S, u64* out, const u64* a, const u64* b, const LimbList& L, int npoly, PolyStride ps, dev_stream st) {
    if (L.n) { const dim3 g = grid3b(S, L.n, npoly, ps); LAUNCH(k_binop<OP_ADD>, g, dim3(TPB), st, S, out, a, b, L, ps); }
}
void launch_sub(KShape S, u64* out, const u64* a, const u64* b, const LimbList& L, int npoly, PolyStride ps, dev_stream st) {
    if (L.n) { const dim3 g = grid3b(S, L.n, npoly, ps); LAUNCH(k_binop<OP_SUB>, g, dim3(TPB), st, S, out, a, b, L, ps); }
}
void launch_mul(KShape S, u64* out, const u64* a, const u64* b, const LimbList& L, int npoly, PolyStride ps, dev_stream st) {
    if (L.n) { const dim3 g = grid3b(S, L.n, npoly, ps); LAUNCH(k_binop<OP_MUL>, g, dim3(TPB), st, S, out, a, b, L, ps); }
}
void launch_neg(KShape S, u64* out, const u64* a, const LimbList& L, int npoly, PolyStride ps, dev_stream st) {
    if (L.n) { const dim3 g = grid3b(S, L.n, npoly, ps); LAUNCH(k_neg, g, dim3(TPB), st, S, out, a, L, ps); }
}
void launch_copy(KShape S, u64* out, const u64* a, int rows, int npoly, PolyStride ps, dev_stream st) {
    if (rows) { const dim3 g = grid3b(S, rows, npoly, ps); LAUNCH(k_copy, g, dim3(TPB), st, S, out, a, ps); }
}
void launch_mul_scalar(KShape S, u64* out, const u64* a, const LimbList& L, const ScalarList& Sc, int npoly, PolyStride ps,
                       dev_stream st) {
    if (L.n) { const dim3 g = grid3b(S, L.n, npoly, ps); LAUNCH(k_mul_scalar, g, dim3(TPB), st, S, out, a, L, Sc, ps); }
}
void launch_sub_mul_scalar(KShape S, u64* out, const u64* a, const u64* b, const LimbList& L, const ScalarList& Sc,
                           int npoly, PolyStride ps, dev_stream st) {
    if (L.n) { const dim3 g = grid3b(S, L.n, npoly, ps); LAUNCH(k_sub_mul_scalar, g, dim3(TPB), st, S, out, a, b, L, Sc, ps); }
}
void launch_mul_const(KShape S, u64* out, const u64* a, const LimbList& L, const ScalarList& CP, const ScalarList& CM,
                      int npoly, PolyStride ps, dev_stream st) {
    if (L.n) { const dim3 g = grid3b(S, L.n, npoly, ps); LAUNCH(k_mul_const<0>, g, dim3(TPB), st, S, out, a, L, CP, CM, ps); }
}
void launch_mac_const(KShape S, u64* acc, const u64* a, const LimbList& L, const ScalarList& CP, const ScalarList& CM,
                      int npoly, PolyStride ps, dev_stream st) {
    if (L.n) { const dim3 g = grid3b(S, L.n, npoly, ps); LAUNCH(k_mul_const<1>, g, dim3(TPB), st, S, acc, a, L, CP, CM, ps); }
}
void launch_add_const(KShape S, u64* out, const u64* a, const LimbList& L, const ScalarList& CP, const ScalarList& CM,
                      int npoly, PolyStride ps, dev_stream st) {
    if (L.n) { const dim3 g = grid3b(S, L.n, npoly, ps); LAUNCH(k_add_const, g, dim3(TPB), st, S, out, a, L, CP, CM, ps, npoly); }
}
void launch_tensor(KShape S, u64* d, const u64* a, const u64* b, const LimbList& L, PolyStride ps, dev_stream st) {
    if (L.n) LAUNCH(k_tensor, grid3(S, L.n, ps.nb > 1 ? ps.nb : 1), dim3(TPB), st, S, d, a, b, L, L.n, ps);
}
void launch_permute(KShape S, u64* out, const u64* a, const u32* perm, int rows, int npoly, PolyStride ps, dev_stream st) {
    if (rows) { const dim3 g = grid3b(S, rows, npoly, ps); LAUNCH(k_permute, g, dim3(TPB), st, S, out, a, perm, ps); }
}
void launch_ks_inner(KShape S, u64* acc, const u64* ext, const u64* own, const u64* evk, const u32* perm,
                     const LimbList& L, const LimbList& ERow, int beta, int evk_rows, int nq, int alpha,
                     const u64* addend, const ScalarList& PmodQ, int accumulate, dev_stream st, bool tensor, KsBatch kb) {
    if (!L.n) return;
    if (beta > KS_MAX_BETA) throw std::runtime_error("ks_inner: too many digits");
    if (kb.nb < 1) kb.nb = 1;
#ifndef KS_INNER_MLP
#define KS_INNER_MLP 1        // 0: the run-time-sized kernel for every call (A/B)
#endif
    if (KS_INNER_MLP && beta <= 3) {
        // digits and items-per-group as template parameters (loads of a whole group in flight together)
#ifndef KS_INNER_NBMAX
#define KS_INNER_NBMAX 2        /* items per group: 2 keeps 64 registers (4 CTAs per SM); 4 needs 125 and measured slower */
#endif
        const int nbg = (kb.nb >= 4 && KS_INNER_NBMAX >= 4) ? 4 : (kb.nb >= 2 && KS_INNER_NBMAX >= 2) ? 2 : 1;
#define KSI(T, NBG, B) LAUNCH((k_ks_inner_mlp<T, NBG, B>), grid3(S, L.n), dim3(TPB), st, S, acc, ext, own, evk, perm, L, ERow, L.n, \
                              evk_rows, nq, alpha, addend, PmodQ, accumulate, kb)
#define KSI_B(T, NBG) do { if (beta == 1) KSI(T, NBG, 1); else if (beta == 2) KSI(T, NBG, 2); else KSI(T, NBG, 3); } while (0)
#define KSI_N(T) do { if (nbg == 4) KSI_B(T, 4); else if (nbg == 2) KSI_B(T, 2); else KSI_B(T, 1); } while (0)
        if (tensor) KSI_N(true); else KSI_N(false);
#undef KSI_N
#undef KSI_B
#undef KSI
        return;
    }
    if (tensor)
        LAUNCH(k_ks_inner<true>, grid3(S, L.n), dim3(TPB), st, S, acc, ext, own, evk, perm, L, ERow, beta, L.n, evk_rows,
               nq, alpha, addend, PmodQ, accumulate, kb);
    else
        LAUNCH(k_ks_inner<false>, grid3(S, L.n), dim3(TPB), st, S, acc, ext, own, evk, perm, L, ERow, beta, L.n, evk_rows,
               nq, alpha, addend, PmodQ, accumulate, kb);
}
void launch_base_convert(KShape S, u64* out, const u64* in, const BaseConvTable* tabs_dev, int tab_zstride, int ns,
                         int max_nt, int nz, size_t in_zs, size_t out_zs, dev_stream st, int mode, int nb, size_t in_bs,
                         size_t out_bs) {
    if (!nz || !max_nt) return;
    if (nb < 1) nb = 1;
    if (mode == BC_FP || mode == BC_FP_GENERIC) {
        dim3 g((1u << S.logn) / TPB, (max_nt + BC_CHUNK - 1) / BC_CHUNK, nz * nb);
#define BCF_CASE(n) case n: \
        if (mode == BC_FP) LAUNCH((k_base_convert_fp<n, 0>), g, dim3(TPB), st, S, out, in, tabs_dev, tab_zstride, in_zs, out_zs, nz, in_bs, out_bs); \
        else LAUNCH((k_base_convert_fp<n, 1>), g, dim3(TPB), st, S, out, in, tabs_dev, tab_zstride, in_zs, out_zs, nz, in_bs, out_bs); \
        return;
        switch (ns) {
            BCF_CASE(1) BCF_CASE(2) BCF_CASE(3) BCF_CASE(4) BCF_CASE(5) BCF_CASE(6) BCF_CASE(7) BCF_CASE(8) BCF_CASE(9) BCF_CASE(10)
            BCF_CASE(11) BCF_CASE(12) BCF_CASE(13) BCF_CASE(14) BCF_CASE(15) BCF_CASE(16)
            default: throw std::runtime_error("base_convert: unsupported source count");
        }
#undef BCF_CASE
    }
#ifndef CKKS_EMU
    if (mode == BC_MMA && ns <= BC_MMA_MAX_SRC) {
        dim3 gm((1u << S.logn) / BCM_TILE, ((max_nt + 7) / 8 + BCM_TG_PER_CTA - 1) / BCM_TG_PER_CTA, nz * nb);
#define BCM_CASE(n) case n: LAUNCH(k_base_convert_mma<n>, gm, dim3(BCM_TILE), st, S, out, in, tabs_dev, tab_zstride, in_zs, out_zs, nz, in_bs, out_bs); return;
        switch (ns) { BCM_CASE(1) BCM_CASE(2) BCM_CASE(3) BCM_CASE(4) BCM_CASE(5) BCM_CASE(6) BCM_CASE(7) BCM_CASE(8) }
#undef BCM_CASE
    }
#endif
    dim3 g((1u << S.logn) / TPB, (max_nt + BC_CHUNK - 1) / BC_CHUNK, nz * nb);
#define BC_CASE(n) case n: LAUNCH(k_base_convert<n>, g, dim3(TPB), st, S, out, in, tabs_dev, tab_zstride, in_zs, out_zs, nz, in_bs, out_bs); break;
    switch (ns) {
        BC_CASE(1) BC_CASE(2) BC_CASE(3) BC_CASE(4) BC_CASE(5) BC_CASE(6) BC_CASE(7) BC_CASE(8) BC_CASE(9) BC_CASE(10)
        BC_CASE(11) BC_CASE(12) BC_CASE(13) BC_CASE(14) BC_CASE(15) BC_CASE(16)
        default: throw std::runtime_error("base_convert: unsupported source count");
    }
#undef BC_CASE
}
void launch_rescale_delta(KShape S, u64* delta, const u64* last, const LimbList& L, int last_mod, int npoly, PolyStride ps,
                          dev_stream st) {
    if (L.n) { const dim3 g = grid3b(S, L.n, npoly, ps); LAUNCH(k_rescale_delta, g, dim3(TPB), st, S, delta, last, L, last_mod, ps); }
}
void launch_center_lift(KShape S, u64* out, const u64* in, const LimbList& L, int src_mod, int npoly, PolyStride ps,
                        dev_stream st) {
    if (L.n) { const dim3 g = grid3b(S, L.n, npoly, ps); LAUNCH(k_center_lift, g, dim3(TPB), st, S, out, in, L, src_mod, ps); }
}
void launch_sample_uniform(KShape S, u64* out, const LimbList& L, u64 seed, u64 stream, dev_stream st) {
    if (L.n) LAUNCH(k_sample_uniform, grid3(S, L.n), dim3(TPB), st, S, out, L, seed, stream);
}
void launch_sample_small(KShape S, u64* out, const LimbList& L, u64 seed, u64 stream, int kind, dev_stream st, int nb,
                         size_t out_bs, const u64* epoch) {
    if (L.n) LAUNCH(k_sample_small, grid3(S, L.n, nb < 1 ? 1 : nb), dim3(TPB), st, S, out, L, seed, stream, kind, out_bs, epoch);
}
void launch_bump(u64* word, dev_stream st) { LAUNCH(k_bump, dim3(1), dim3(32), st, word); }
void launch_reduce_i64(KShape S, u64* out, const i64* v, const LimbList& L, dev_stream st, int nb, size_t out_bs) {
    if (L.n) LAUNCH(k_reduce_i64, grid3(S, L.n, nb < 1 ? 1 : nb), dim3(TPB), st, S, out, v, L, out_bs);
}
// decode direction: w (n complex, natural order) -> z in `out`
#ifndef FFT_FUSED
#define FFT_FUSED 1            // 0: one launch per butterfly stage (round-1 form, kept for A/B and as the reference form)
#endif
void launch_special_fft(KShape S, double* out, const double* w, const u32* rot, const double* ksi, dev_stream st, int nb) {
    const int n = 1 << (S.logn - 1);
    if (FFT_FUSED && (n == FFT_LOW || n == FFT_LOW * FFT_HI_MAX)) {
        LAUNCH(k_fft_low, dim3(n / FFT_LOW, nb), dim3(FFT_LOW / 2), st, S, out, w, rot, ksi, 0, 1.0);
        if (n > FFT_LOW) LAUNCH((k_fft_high<FFT_HI_MAX, 0>), dim3(FFT_LOW / TPB, nb), dim3(TPB), st, S, out, rot, ksi);
        return;
    }
    LAUNCH(k_bitrev_copy, dim3(n / TPB, nb), dim3(TPB), st, S, out, w, 1.0);
    for (int len = 2; len <= n; len <<= 1)
        LAUNCH(k_fft_stage, dim3(n / 2 / TPB, nb), dim3(TPB), st, S, out, rot, ksi, len, 0);
}
// encode direction: z -> w (includes 1/n); `z` is overwritten as scratch
void launch_special_ifft(KShape S, double* out, double* z, const u32* rot, const double* ksi, dev_stream st, int nb) {
    const int n = 1 << (S.logn - 1);
    if (FFT_FUSED && (n == FFT_LOW || n == FFT_LOW * FFT_HI_MAX)) {
        if (n > FFT_LOW) LAUNCH((k_fft_high<FFT_HI_MAX, 1>), dim3(FFT_LOW / TPB, nb), dim3(TPB), st, S, z, rot, ksi);
        LAUNCH(k_fft_low, dim3(n / FFT_LOW, nb), dim3(FFT_LOW / 2), st, S, out, z, rot, ksi, 1, 1.0 / (double)n);
        return;
    }
    for (int len = n; len >= 2; len >>= 1)
        LAUNCH(k_fft_stage, dim3(n / 2 / TPB, nb), dim3(TPB), st, S, z, rot, ksi, len, 1);
    LAUNCH(k_bitrev_copy, dim3(n / TPB, nb), dim3(TPB), st, S, out, z, 1.0 / (double)n);
}
void launch_snap_zeta16(KShape S, double* z, const double* table, int stride, dev_stream st, int nb) {
    LAUNCH(k_snap_zeta16, dim3((1u << (S.logn - 1)) / TPB, nb), dim3(TPB), st, S, z, table, stride);
}
void launch_zeta16_from_nibbles(KShape S, double* z, const unsigned char* nib, const double* table, dev_stream st, int nb) {
    LAUNCH(k_zeta16_from_nibbles, dim3((1u << (S.logn - 1)) / TPB, nb), dim3(TPB), st, S, z, nib, table);
}
void launch_nibbles_from_zeta16(KShape S, unsigned char* nib, const double* z, dev_stream st, int nb) {
    LAUNCH(k_nibbles_from_zeta16, dim3((1u << (S.logn - 1)) / TPB, nb), dim3(TPB), st, S, nib, z);
}
void launch_round_coeffs(KShape S, i64* out, const double* w, double scale, int* flag, dev_stream st, int nb) {
    LAUNCH(k_round_coeffs, dim3((1u << (S.logn - 1)) / TPB, nb), dim3(TPB), st, S, out, w, scale, flag);
}
void launch_center_to_w(KShape S, double* w, const u64* coef, int mod, double scale, dev_stream st, int nb) {
    LAUNCH(k_center_to_w, dim3((1u << (S.logn - 1)) / TPB, nb), dim3(TPB), st, S, w, coef, mod, scale);
}
