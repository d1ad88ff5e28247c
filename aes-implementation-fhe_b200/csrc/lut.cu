// lut.cu -- fused sparse multiply-accumulate over ciphertext power bases (the LUT polynomials of the
// Zeta16 encoding: reference xor4_lut.py:63-74, mixcol_final.py:80-99, invmixcolumns_fhe.py:76-90,
// sub_bytes_lut.py:63-71).
//
//   lut2:    sum_t c_t * A[p_t] (x) B[q_t]   as ONE 3-polynomial accumulation and ONE relinearisation whose ModDown
//            also divides by the two rescale primes, instead of one ct*ct multiplication (key switch + rescale) per term.  Terms are
//            grouped by p:  sum_p A_p (x) (sum_q c_pq B_q), so the kernel does |P| tensor products.
//   lincomb: sum_k c_k * X[k] with one rescale per distinct input level (not one per term).
//
// Constants follow spec S7 (two scalars per limb).  Algorithmic bytes of k_lut2:
// 2(|P|+|Q|)(l+1) N w read + 3(l+1) N w written.
#include <algorithm>

#include "engine.cuh"

#define LUT_MAX_BASIS 16
#ifndef LUT_FP
#define LUT_FP 1              // rows of FP64 moduli (below CKKS_FP_LIMIT) of the LUT / BSGS multiply-accumulates on the FP64 pipe
#endif
#define LINCOMB_MAX 128

struct Lut2Args {
    const u64* a[LUT_MAX_BASIS];       // basis ciphertexts, all at the same level: [2][rows][N]
    const u64* b[LUT_MAX_BASIS];
    const unsigned char* tp;           // [nterms] p index, sorted by p
    const unsigned char* tq;           // [nterms]
    const u64* consts;                 // [rows][2 halves][nterms][2] = {c, shoup(c)}
    int nterms, rows;
    unsigned abm, bbm;                 // bit p set: basis element p is batched (advances by 2 rows N per batch item = blockIdx.z)
};
struct LinCombArgs {
    const u64* x[LINCOMB_MAX];         // [npoly][rows][N] each (same level)
    const u64* consts;                 // [rows][2][nterms][2]
    int nterms, rows;
    int npoly;                         // blockIdx.z = batch item * npoly + polynomial
    u64 xbm[LINCOMB_MAX / 64];         // bit t set: x[t] is batched
};

struct DiagMacArgs {
    const u64* x[16];                  // ciphertexts [nb][2][rows][N] (blockIdx.z = batch item * 2 + polynomial)
    const u64* p[16];                  // plaintext polynomials [rows][N], shared by the batch
    int nterms, rows;
};

#define DIAG_ROWS_MAX 8
struct DiagRowsArgs {
    const u64* x[16];                  // baby-step ciphertexts [nb][2][rows][N]
    const u64* p[DIAG_ROWS_MAX][16];   // p[r][t]: diagonal of giant row r that multiplies baby t (null: none)
    u64* out[DIAG_ROWS_MAX];           // inner sum of giant row r, [nb][2][rows][N]
    int nx, nrows, rows, nz;           // nz = 2 * nb slices (batch item * 2 + polynomial)
};

namespace {

__device__ __forceinline__ double bits2dbl(u64 x) { union { double d; u64 u; } c; c.u = x; return c.d; }

// All inner sums of one BSGS matrix in one pass: out_r = sum_t x_t (.) p[r][t] for every giant row r.  A baby-step word is
// loaded ONCE and multiplied into every row that uses it (the one-row kernel re-reads all babies per row: 8 x the traffic of
// a radix-32 matrix); the nz slices of a coefficient block sit next to each other in the launch order, so the diagonals
// (shared by the batch) come from L2 for all but the first of them.
template <int NR>
__global__ void __launch_bounds__(256, 4)
k_diag_mac_rows(KShape S, const GRID_CONST DiagRowsArgs G, const GRID_CONST LimbList L) {
    const int row = blockIdx.y;
    const ModConst m = S.mc[L.idx[row]];
    const size_t N = (size_t)1 << S.logn, P = (size_t)G.rows << S.logn;
    const unsigned cb = blockIdx.x / (unsigned)G.nz, z = blockIdx.x - cb * (unsigned)G.nz;
    if (LUT_FP && m.q < CKKS_FP_LIMIT) {
        // rows of FP64 moduli (every row of the default chain): exact modular products on the FP64 pipe, |r| <= 0.91 q each,
        // folded every 4 babies so that a sum stays below 2^53 (the 128-bit integer multiply-accumulate below is 3 x slower)
        const double qd = ull2d_rn(m.q), qinv = fdiv_rn(1.0, qd);
        FOR_THREADS {
            const size_t i = (size_t)row * N + cb * 256 + threadIdx.x;
            double acc[NR];
#pragma unroll
            for (int r = 0; r < NR; r++) acc[r] = 0.0;
#pragma unroll 2
            for (int t = 0; t < G.nx; t++) {
                const double xv = ull2d_rn(ldg(G.x[t] + z * P + i));
                u64 pv[NR];
#pragma unroll
                for (int r = 0; r < NR; r++) {
                    const u64* pp = G.p[r][t];
                    pv[r] = pp ? ldg(pp + i) : 0;
                }
#pragma unroll
                for (int r = 0; r < NR; r++) acc[r] = fadd_rn(acc[r], modmul_fp(xv, ull2d_rn(pv[r]), qd, qinv));
                if ((t & 3) == 3 && t + 1 < G.nx) {
#pragma unroll
                    for (int r = 0; r < NR; r++) acc[r] = fold_fp(acc[r], qd, qinv);
                }
            }
#pragma unroll
            for (int r = 0; r < NR; r++)
                if (r < G.nrows) G.out[r][z * P + i] = canon_fp(acc[r], qd, qinv);
        }
        return;
    }
    FOR_THREADS {
        const size_t i = (size_t)row * N + cb * 256 + threadIdx.x;
        u64 hi[NR], lo[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) hi[r] = lo[r] = 0;
        for (int t = 0; t < G.nx; t++) {
            const u64 xv = ldg(G.x[t] + z * P + i);
            u64 pv[NR];
#pragma unroll
            for (int r = 0; r < NR; r++) {             // the row's diagonals are requested together, then multiplied
                const u64* pp = G.p[r][t];
                pv[r] = pp ? ldg(pp + i) : 0;
            }
#pragma unroll
            for (int r = 0; r < NR; r++) mac128(hi[r], lo[r], xv, pv[r]);
            if ((t & 3) == 3 && t + 1 < G.nx) {        // a row takes at most one product per baby: sums stay below 4 q^2
#pragma unroll
                for (int r = 0; r < NR; r++) { lo[r] = barrett_reduce128(hi[r], lo[r], m); hi[r] = 0; }
            }
        }
#pragma unroll
        for (int r = 0; r < NR; r++)
            if (r < G.nrows) G.out[r][z * P + i] = barrett_reduce128(hi[r], lo[r], m);
    }
}

// out[poly][rows][N] = sum_t x_t[poly] * p_t   (BSGS inner sum of a linear transform, un-rescaled; blockIdx.z = poly)
__global__ void __launch_bounds__(256)
k_diag_mac(KShape S, u64* __restrict__ out, const GRID_CONST DiagMacArgs G, const GRID_CONST LimbList L) {
    const int row = blockIdx.y;
    const ModConst m = S.mc[L.idx[row]];
    const size_t N = (size_t)1 << S.logn, P = (size_t)G.rows << S.logn;
    FOR_THREADS {
        const size_t i = (size_t)row * N + blockIdx.x * 256 + threadIdx.x;
        u64 hi = 0, lo = 0;
        for (int t = 0; t < G.nterms; t++) {
            mac128(hi, lo, ldg(G.x[t] + blockIdx.z * P + i), ldg(G.p[t] + i));
            if ((t & 3) == 3 && t + 1 < G.nterms) { lo = barrett_reduce128(hi, lo, m); hi = 0; }
        }
        out[blockIdx.z * P + i] = barrett_reduce128(hi, lo, m);
    }
}

// d[3][rows][N] = sum over p-groups of A_p (x) (sum_q c_pq B_q)
__global__ void __launch_bounds__(256)
k_lut2(KShape S, u64* __restrict__ d, const GRID_CONST Lut2Args G, const GRID_CONST LimbList L) {
    const int row = blockIdx.y;
    const ModConst m = S.mc[L.idx[row]];
    const size_t N = (size_t)1 << S.logn, P = (size_t)G.rows << S.logn;
    const int half = blockIdx.x < (gridDim.x >> 1) ? 0 : 1;
    const u64* cst = G.consts + ((size_t)(row * 2 + half) * G.nterms) * 2;
    const size_t zb = (size_t)blockIdx.z * 2 * P;       // offset of this batch item inside a batched basis element
    d += (size_t)blockIdx.z * 3 * P;
    if (LUT_FP && m.q < CKKS_FP_LIMIT) {
        // FP64 rows: the inner sums u = sum_q c_pq B_q and the tensor products A_p (x) u as exact FP64 modular products
        // (|r| <= 0.91 q), sums folded every 4 terms / 4 groups; u is folded to |u| <= q/2 before it multiplies
        const double qd = ull2d_rn(m.q), qinv = fdiv_rn(1.0, qd);
        FOR_THREADS {
            const size_t i = (size_t)row * N + blockIdx.x * 256 + threadIdx.x;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            int t = 0, groups = 0;
            while (t < G.nterms) {
                const int p = ldg(G.tp + t);
                double u0 = 0.0, u1 = 0.0;
                int cnt = 0;
                for (; t < G.nterms && ldg(G.tp + t) == p; t++) {
                    const int q = ldg(G.tq + t);
                    const u64* bq = G.b[q] + (((G.bbm >> q) & 1u) ? zb : 0);
                    const double c = bits2dbl(ldg(cst + 2 * t + 1));
                    u0 = fadd_rn(u0, modmul_fp(ull2d_rn(ldg(bq + i)), c, qd, qinv));
                    u1 = fadd_rn(u1, modmul_fp(ull2d_rn(ldg(bq + i + P)), c, qd, qinv));
                    if ((++cnt & 3) == 0) { u0 = fold_fp(u0, qd, qinv); u1 = fold_fp(u1, qd, qinv); }
                }
                u0 = fold_fp(u0, qd, qinv);
                u1 = fold_fp(u1, qd, qinv);
                const u64* ap = G.a[p] + (((G.abm >> p) & 1u) ? zb : 0);
                const double a0 = ull2d_rn(ldg(ap + i)), a1 = ull2d_rn(ldg(ap + i + P));
                s0 = fadd_rn(s0, modmul_fp(a0, u0, qd, qinv));
                s1 = fadd_rn(s1, fadd_rn(modmul_fp(a0, u1, qd, qinv), modmul_fp(a1, u0, qd, qinv)));
                s2 = fadd_rn(s2, modmul_fp(a1, u1, qd, qinv));
                if ((++groups & 1) == 0) {             // s1 takes two products per group: fold every 2 groups
                    s0 = fold_fp(s0, qd, qinv); s1 = fold_fp(s1, qd, qinv); s2 = fold_fp(s2, qd, qinv);
                }
            }
            d[i] = canon_fp(s0, qd, qinv);
            d[i + P] = canon_fp(s1, qd, qinv);
            d[i + 2 * P] = canon_fp(s2, qd, qinv);
        }
        return;
    }
    FOR_THREADS {
        const size_t i = (size_t)row * N + blockIdx.x * 256 + threadIdx.x;
        u64 h0 = 0, l0 = 0, h1 = 0, l1 = 0, h2 = 0, l2 = 0;
        int t = 0, groups = 0;
        while (t < G.nterms) {
            const int p = ldg(G.tp + t);
            u64 u0 = 0, u1 = 0;
            int cnt = 0;
            for (; t < G.nterms && ldg(G.tp + t) == p; t++) {
                const int q = ldg(G.tq + t);
                const u64* bq = G.b[q] + (((G.bbm >> q) & 1u) ? zb : 0);
                const u64 c = ldg(cst + 2 * t), cs = ldg(cst + 2 * t + 1);
                u0 += shoup_mul(ldg(bq + i), c, cs, m.q);
                u1 += shoup_mul(ldg(bq + i + P), c, cs, m.q);
                if ((++cnt & 7) == 0) { u0 = barrett_reduce64(u0, m); u1 = barrett_reduce64(u1, m); }
            }
            u0 = barrett_reduce64(u0, m);
            u1 = barrett_reduce64(u1, m);
            const u64* ap = G.a[p] + (((G.abm >> p) & 1u) ? zb : 0);
            const u64 a0 = ldg(ap + i), a1 = ldg(ap + i + P);
            mac128(h0, l0, a0, u0);
            mac128(h1, l1, a0, u1);
            mac128(h1, l1, a1, u0);
            mac128(h2, l2, a1, u1);
            if ((++groups & 1) == 0) {                 // keep the 128-bit sums below 4 q^2
                l0 = barrett_reduce128(h0, l0, m); h0 = 0;
                l1 = barrett_reduce128(h1, l1, m); h1 = 0;
                l2 = barrett_reduce128(h2, l2, m); h2 = 0;
            }
        }
        d[i] = barrett_reduce128(h0, l0, m);
        d[i + P] = barrett_reduce128(h1, l1, m);
        d[i + 2 * P] = barrett_reduce128(h2, l2, m);
    }
}

// out[poly][rows][N] = sum_t x_t * c_t  (blockIdx.z = polynomial)
__global__ void __launch_bounds__(256)
k_lincomb(KShape S, u64* __restrict__ out, const GRID_CONST LinCombArgs G, const GRID_CONST LimbList L) {
    const int row = blockIdx.y;
    const ModConst m = S.mc[L.idx[row]];
    const size_t N = (size_t)1 << S.logn, P = (size_t)G.rows << S.logn;
    const int half = blockIdx.x < (gridDim.x >> 1) ? 0 : 1;
    const u64* cst = G.consts + ((size_t)(row * 2 + half) * G.nterms) * 2;
    const unsigned bi = blockIdx.z / (unsigned)G.npoly, kp = blockIdx.z - bi * (unsigned)G.npoly;
    const size_t zb = (size_t)bi * G.npoly * P;          // offset of this batch item inside a batched operand
    if (LUT_FP && m.q < CKKS_FP_LIMIT) {
        const double qd = ull2d_rn(m.q), qinv = fdiv_rn(1.0, qd);
        FOR_THREADS {
            const size_t i = kp * P + (size_t)row * N + blockIdx.x * 256 + threadIdx.x;
            double acc = 0.0;
            for (int t = 0; t < G.nterms; t++) {
                const u64* xt = G.x[t] + (((G.xbm[t >> 6] >> (t & 63)) & 1ull) ? zb : 0);
                acc = fadd_rn(acc, modmul_fp(ull2d_rn(ldg(xt + i)), bits2dbl(ldg(cst + 2 * t + 1)), qd, qinv));
                if ((t & 3) == 3) acc = fold_fp(acc, qd, qinv);
            }
            out[zb + i] = canon_fp(acc, qd, qinv);
        }
        return;
    }
    FOR_THREADS {
        const size_t i = kp * P + (size_t)row * N + blockIdx.x * 256 + threadIdx.x;
        u64 acc = 0;
        for (int t = 0; t < G.nterms; t++) {
            const u64* xt = G.x[t] + (((G.xbm[t >> 6] >> (t & 63)) & 1ull) ? zb : 0);
            acc += shoup_mul(ldg(xt + i), ldg(cst + 2 * t), ldg(cst + 2 * t + 1), m.q);
            if ((t & 7) == 7) acc = barrett_reduce64(acc, m);
        }
        out[zb + i] = barrett_reduce64(acc, m);
    }
}

}  // namespace

namespace ckks {

// device table of constants {c, shoup(c)} laid out [row][half][term]; cached by content
const u64* Engine::const_table(const double* coef_re_im, int n, int scale_level, int level) {
    // FNV-1a over the coefficient bytes + shape
    u64 h = 1469598103934665603ull;
    const unsigned char* bytes = reinterpret_cast<const unsigned char*>(coef_re_im);
    for (size_t i = 0; i < (size_t)n * 16; i++) { h ^= bytes[i]; h *= 1099511628211ull; }
    h ^= (u64)n * 0x9E3779B97F4A7C15ull + ((u64)scale_level << 20) + ((u64)level << 8);
    auto it = const_tabs.find(h);
    if (it != const_tabs.end()) return it->second;
    const int rows = level + 1;
    std::vector<int> idx = mods_q(level);
    std::vector<u64> host((size_t)rows * 2 * n * 2);
    ScalarList cp, cm;
    for (int t = 0; t < n; t++) {
        const_residues(coef_re_im[2 * t], coef_re_im[2 * t + 1], scales[scale_level], idx, cp, cm);
        for (int r = 0; r < rows; r++) {
            u64* p0 = &host[(((size_t)r * 2 + 0) * n + t) * 2];
            u64* p1 = &host[(((size_t)r * 2 + 1) * n + t) * 2];
            p0[0] = cp.v[r]; p0[1] = cp.vs[r];
            p1[0] = cm.v[r]; p1[1] = cm.vs[r];
            if (LUT_FP && mod[idx[r]] < CKKS_FP_LIMIT) {
                // FP64 rows read the constant as a double (no conversion per term in the kernels): the companion slot
                // holds its bits instead of the Shoup word the integer rows use
                const double d0 = (double)cp.v[r], d1 = (double)cm.v[r];
                memcpy(&p0[1], &d0, 8);
                memcpy(&p1[1], &d1, 8);
            }
        }
    }
    u64* d = alloc(host.size());
    dev::h2d(d, host.data(), host.size() * sizeof(u64), st);
    dev::sync(st);
    const_tabs[h] = d;
    return d;
}

Ct* Engine::lut2(const std::vector<Ct*>& A, const std::vector<Ct*>& B, const int* p, const int* q, const double* coef,
                 int nterms) {
    if (!has_relin) throw std::runtime_error("lut2 needs a relinearisation key");
    if (nterms < 1) throw std::runtime_error("lut2: empty term list");
    if ((int)A.size() > LUT_MAX_BASIS || A.size() != B.size()) throw std::runtime_error("lut2: bad basis size");
    // sort terms by p (stable), find the common level
    std::vector<int> order(nterms);
    for (int t = 0; t < nterms; t++) order[t] = t;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return p[x] < p[y]; });
    int level = L();
    for (int t = 0; t < nterms; t++) {
        if (p[t] < 0 || p[t] >= (int)A.size() || q[t] < 0 || q[t] >= (int)B.size() || !A[p[t]] || !B[q[t]])
            throw std::runtime_error("lut2: term refers to a missing basis element");
        if (A[p[t]]->npoly != 2 || B[q[t]]->npoly != 2) throw PolyCountError("lut2: operands should have 2 polynomials");
        level = std::min(level, std::min(A[p[t]]->level, B[q[t]]->level));
    }
    need_levels(level, 2, "lut2");
    Lut2Args G;
    memset(&G, 0, sizeof(G));
    std::vector<unsigned char> tp(nterms), tq(nterms);
    std::vector<double> cs(2 * (size_t)nterms);
    int nb = 1;
    for (int t = 0; t < nterms; t++) {
        const int s = order[t];
        tp[t] = (unsigned char)p[s];
        tq[t] = (unsigned char)q[s];
        cs[2 * t] = coef[2 * s];
        cs[2 * t + 1] = coef[2 * s + 1];
        G.a[p[s]] = level_down(A[p[s]], level)->d;
        G.b[q[s]] = level_down(B[q[s]], level)->d;
        // an nb = 1 basis (a round key's) is broadcast to the batch of the other one
        for (const Ct* c : {A[p[s]], B[q[s]]}) {
            if (c->nb != 1 && nb != 1 && c->nb != nb) throw std::runtime_error("lut2: basis elements hold different batch sizes");
            nb = std::max(nb, c->nb);
        }
        if (A[p[s]]->nb > 1) G.abm |= 1u << p[s];
        if (B[q[s]]->nb > 1) G.bbm |= 1u << q[s];
    }
    // constants at scale S[level-1]: S_l^2 * S_{l-1} / (q_l q_{l-1}) = S_{l-2}
    G.consts = const_table(cs.data(), nterms, level - 1, level);
    // term index lists: cached next to the constants (same content hash + tag)
    std::vector<double> tagged(2 * (size_t)nterms);
    for (int t = 0; t < nterms; t++) { tagged[2 * t] = tp[t]; tagged[2 * t + 1] = tq[t]; }
    u64 h = 1469598103934665603ull;
    for (int t = 0; t < nterms; t++) { h ^= tp[t] + 256u * tq[t] + 65536u * (u64)t; h *= 1099511628211ull; }
    auto it = index_tabs.find(h);
    if (it == index_tabs.end()) {
        unsigned char* dtab = (unsigned char*)dev::alloc(2 * (size_t)nterms, st);
        dev::h2d(dtab, tp.data(), nterms, st);
        dev::h2d(dtab + nterms, tq.data(), nterms, st);
        dev::sync(st);
        it = index_tabs.emplace(h, dtab).first;
    }
    G.tp = it->second;
    G.tq = it->second + nterms;
    G.nterms = nterms;
    G.rows = level + 1;
    const size_t n = N(), ps = (size_t)(level + 1) * n;
    LimbList ll = limb_list(mods_q(level));
    u64* d = alloc((size_t)nb * 3 * ps);
    LAUNCH(k_lut2, dim3((unsigned)(n / 256), level + 1, nb), dim3(256), st, ks, d, G, ll);
    // ONE relinearisation merged with BOTH rescales: (<digits(d2), rlk> + P (d0, d1)) / (P q_l q_{l-1})   (spec S6b)
    Decomp D = decompose(d + 2 * ps, level, nullptr, nb, 3 * ps, 0);
    Ct* out = new_ct(2, level - 2, nb);
    ks_apply(D, &relin, nullptr, out->d, d, 2, false, 3 * ps);
    release(D.ext);
    release(d);
    n_mul_cc += nb;
    return out;
}

// out = sum_t x_t (.) p_t for up to 16 (ciphertext, plaintext) pairs at the same level, no rescale
void Engine::diag_mac(u64* out, const std::vector<const Ct*>& x, const std::vector<const Pt*>& p, int level, int nb) {
    if (x.empty() || x.size() != p.size() || x.size() > 16) throw std::runtime_error("diag_mac: 1..16 terms");
    DiagMacArgs G;
    memset(&G, 0, sizeof(G));
    for (size_t t = 0; t < x.size(); t++) {
        if (x[t]->level != level || p[t]->level != level || x[t]->npoly != 2 || x[t]->nb != nb)
            throw std::runtime_error("diag_mac: level mismatch");
        G.x[t] = x[t]->d;
        G.p[t] = p[t]->d;
    }
    G.nterms = (int)x.size();
    G.rows = level + 1;
    // [nb][2] slices with uniform strides: blockIdx.z = batch item * 2 + polynomial
    LAUNCH(k_diag_mac, dim3((unsigned)(N() / 256), level + 1, 2 * nb), dim3(256), st, ks, out, G, limb_list(mods_q(level)));
}

// every inner sum of a BSGS matrix in one launch: out[r] = sum_t x[t] (.) p[r][t] (p[r][t] null: baby t is not in row r)
void Engine::diag_mac_rows(const std::vector<u64*>& out, const std::vector<const Ct*>& x,
                           const std::vector<std::vector<const Pt*>>& p, int level, int nb) {
    if (x.empty() || x.size() > 16 || out.empty() || out.size() > DIAG_ROWS_MAX || p.size() != out.size())
        throw std::runtime_error("diag_mac_rows: 1..16 babies, 1..8 rows");
    DiagRowsArgs G;
    memset(&G, 0, sizeof(G));
    for (size_t t = 0; t < x.size(); t++) {
        if (x[t]->level != level || x[t]->npoly != 2 || x[t]->nb != nb) throw std::runtime_error("diag_mac_rows: level mismatch");
        G.x[t] = x[t]->d;
    }
    for (size_t r = 0; r < out.size(); r++) {
        if (p[r].size() != x.size()) throw std::runtime_error("diag_mac_rows: one diagonal slot per baby");
        for (size_t t = 0; t < x.size(); t++) {
            if (p[r][t] && p[r][t]->level != level) throw std::runtime_error("diag_mac_rows: level mismatch");
            G.p[r][t] = p[r][t] ? p[r][t]->d : nullptr;
        }
        G.out[r] = out[r];
    }
    G.nx = (int)x.size();
    G.nrows = (int)out.size();
    G.rows = level + 1;
    G.nz = 2 * nb;
    const dim3 grid((unsigned)(N() / 256) * (unsigned)G.nz, level + 1, 1);
    if (G.nrows <= 4) LAUNCH(k_diag_mac_rows<4>, grid, dim3(256), st, ks, G, limb_list(mods_q(level)));
    else LAUNCH(k_diag_mac_rows<DIAG_ROWS_MAX>, grid, dim3(256), st, ks, G, limb_list(mods_q(level)));
}

// sum_k c_k X_k: one fused multiply-accumulate and one rescale per distinct input level, partial sums added
// from the highest level down (each addition aligns the running sum by one level_down)
Ct* Engine::lincomb(const std::vector<Ct*>& X, const double* coef, int n) {
    if (n < 1) throw std::runtime_error("lincomb: empty");
    std::map<int, std::vector<int>, std::greater<int>> by_level;
    if (!X[0]) throw std::runtime_error("lincomb: missing ciphertext");
    int npoly = X[0]->npoly, nb = 1;
    for (int k = 0; k < n; k++) {
        if (!X[k]) throw std::runtime_error("lincomb: missing ciphertext");
        if (X[k]->npoly != npoly) throw PolyCountError("lincomb: mixed polynomial counts");
        if (X[k]->nb != 1 && nb != 1 && X[k]->nb != nb) throw std::runtime_error("lincomb: operands hold different batch sizes");
        nb = std::max(nb, X[k]->nb);
        need_levels(X[k]->level, 1, "lincomb");
        by_level[X[k]->level].push_back(k);
    }
    Ct* acc = nullptr;
    for (auto& kv : by_level) {
        const int level = kv.first;
        const size_t nn = N(), ps = (size_t)(level + 1) * nn;
        LimbList ll = limb_list(mods_q(level));
        u64* sum = nullptr;                                  // un-rescaled partial sum at this level
        for (size_t off = 0; off < kv.second.size(); off += LINCOMB_MAX) {
            const int cnt = (int)std::min((size_t)LINCOMB_MAX, kv.second.size() - off);
            LinCombArgs G;
            memset(&G, 0, sizeof(G));
            std::vector<double> cs(2 * (size_t)cnt);
            for (int t = 0; t < cnt; t++) {
                const int k = kv.second[off + t];
                G.x[t] = X[k]->d;
                if (X[k]->nb > 1) G.xbm[t >> 6] |= 1ull << (t & 63);
                cs[2 * t] = coef[2 * k];
                cs[2 * t + 1] = coef[2 * k + 1];
            }
            G.consts = const_table(cs.data(), cnt, level, level);
            G.nterms = cnt;
            G.rows = level + 1;
            G.npoly = npoly;
            u64* part = alloc((size_t)nb * npoly * ps);
            LAUNCH(k_lincomb, dim3((unsigned)(nn / 256), level + 1, npoly * nb), dim3(256), st, ks, part, G, ll);
            if (!sum) sum = part;
            else {
                launch_add(ks, sum, sum, part, ll, npoly * nb, PolyStride{ps, ps, ps}, st);
                release(part);
            }
        }
        Ct* r = new_ct(npoly, level - 1, nb);
        rescale_into(r->d, sum, npoly, level, nb);
        release(sum);
        if (!acc) acc = r;
        else {
            Ct* s = add(acc, r);
            free_ct(acc);
            free_ct(r);
            acc = s;
        }
    }
    return acc;
}

}  // namespace ckks
