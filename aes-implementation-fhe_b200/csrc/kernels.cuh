// kernels.cuh -- launch wrappers for the streaming RNS kernels (see kernels.cu).
#pragma once
#include "common.cuh"

#define BC_MAX_SRC 16
#define BC_MMA_MAX_SRC 8
#define BC_MAX_TGT 48
#ifndef BC_CHUNK
#define BC_CHUNK 24
#endif

// element strides between consecutive polynomials for each pointer of a kernel.  Batched ciphertexts (engine.cuh: Ct::nb
// independent ciphertexts of one shape in one buffer, [nb][npoly][rows][N]) add a second set of strides between
// consecutive batch items; blockIdx.z then runs over nb * npoly (batch-major).  A stride of 0 broadcasts an operand
// (a plaintext, a key, an nb = 1 ciphertext) to every polynomial / batch item.
struct PolyStride {
    size_t out, a, b;
    int nb = 1;                       // batch items
    size_t bout = 0, ba = 0, bb = 0;  // element strides between consecutive batch items
    int npoly = 0;                    // set by the launch wrapper: polynomials per batch item (0: blockIdx.z is the polynomial)
};
// fast basis conversion {src} -> {tgt}; hat is a device array [ns][nt]
struct BaseConvTable {
    int ns, nt;
    unsigned char src[BC_MAX_SRC];     // modulus index of source row i
    unsigned char srow[BC_MAX_SRC];    // input row of source i
    unsigned char tgt[BC_MAX_TGT];     // modulus index of target t
    unsigned char orow[BC_MAX_TGT];    // output row of target t
    u64 hatinv[BC_MAX_SRC], hatinv_s[BC_MAX_SRC];
    const u64* hat;
    // exact mode (ModDown): the overflow count u = round(sum_i y_i / s_i) is computed in fp64 and u * D is taken out,
    // so the conversion returns the CENTRED residue of the input modulo D = prod s_i (spec S5')
    int exact;
    double inv_src[BC_MAX_SRC];        // 1 / s_i
    u64 negD[BC_MAX_TGT];              // q_t - (D mod q_t)
    // tensor-core path (ns <= 8, k_base_convert_mma): the hat matrix byte-sliced and laid out in mma.m16n8k32 A-fragment
    // order, [target group of 8][diagonal pair j][k-step][lane] x 16 bytes, and 2^(8 d) mod q_t, [nt][16]
    const void* afrag;
    const u64* pow8;
    // FP64 path (k_base_convert_fp): a target below CKKS_FP_LIMIT may be evaluated on the FP64 pipe as a sum of exact
    // FP64 modular products; a source modulus at or above CKKS_FP_LIMIT (q_0, 61-bit special primes) enters as two halves below 2^32
    unsigned char tfp[BC_MAX_TGT];     // 1: target t takes the FP64 path
    u32 swide;                         // bit i: source i is split at 2^32
    const double* hatf;                // [nt][ns][2]: hat, hat 2^32 mod q_t  (FP64 targets)
    double tqinv[BC_MAX_TGT];          // 1 / q_t
    double negDd[BC_MAX_TGT];          // negD as a double
};
// how launch_base_convert evaluates the sums
// BC_FP: only source 0 of a table may be wide (branch-free source loop); BC_FP_GENERIC: any source (run-time mask)
enum BcMode { BC_INT = 0, BC_MMA = 1, BC_FP = 2, BC_FP_GENERIC = 3 };
// shape shared by all kernels: N coefficients per row, modulus table
struct KShape {
    const ModConst* mc;
    int logn;
};

void launch_add(KShape S, u64* out, const u64* a, const u64* b, const LimbList& L, int npoly, PolyStride ps, dev_stream st);
void launch_sub(KShape S, u64* out, const u64* a, const u64* b, const LimbList& L, int npoly, PolyStride ps, dev_stream st);
void launch_mul(KShape S, u64* out, const u64* a, const u64* b, const LimbList& L, int npoly, PolyStride ps, dev_stream st);
void launch_neg(KShape S, u64* out, const u64* a, const LimbList& L, int npoly, PolyStride ps, dev_stream st);
void launch_mul_scalar(KShape S, u64* out, const u64* a, const LimbList& L, const ScalarList& Sc, int npoly, PolyStride ps, dev_stream st);
void launch_sub_mul_scalar(KShape S, u64* out, const u64* a, const u64* b, const LimbList& L, const ScalarList& Sc, int npoly, PolyStride ps, dev_stream st);
void launch_mul_const(KShape S, u64* out, const u64* a, const LimbList& L, const ScalarList& CP, const ScalarList& CM, int npoly, PolyStride ps, dev_stream st);
void launch_mac_const(KShape S, u64* acc, const u64* a, const LimbList& L, const ScalarList& CP, const ScalarList& CM, int npoly, PolyStride ps, dev_stream st);
// out = a with the constant added to polynomial 0 of every batch item; the other polynomials are copied
void launch_add_const(KShape S, u64* out, const u64* a, const LimbList& L, const ScalarList& CP, const ScalarList& CM, int npoly, PolyStride ps, dev_stream st);
// d[nb][3][nl][N] from a[nb|1][2][nl][N], b[nb|1][2][nl][N]; ps.nb / bout / ba / bb give the batch layout (poly strides unused)
void launch_tensor(KShape S, u64* d, const u64* a, const u64* b, const LimbList& L, PolyStride ps, dev_stream st);
void launch_copy(KShape S, u64* out, const u64* a, int rows, int npoly, PolyStride ps, dev_stream st);
void launch_permute(KShape S, u64* out, const u64* a, const u32* perm, int rows, int npoly, PolyStride ps, dev_stream st);
// batch layout of a key-switch inner product: element strides between consecutive batch items (0 = shared operand)
struct KsBatch {
    int nb = 1;
    size_t acc = 0, ext = 0, own = 0, addend = 0;
    // non-tensor addend: 0 = both polynomials ([2][nq][N]); 1 = polynomial 0 only (the c0 of a Galois map: c0' = sigma(c0)
    // + ks0, the addition that used to be a launch of its own); 2 = polynomial 0 only, gathered through `perm` like the
    // digits (hoisted rotations: the un-permuted input ciphertext is the addend, no permutation launch either)
    int addend_mode = 0;
};
// acc[2][rows][N] = sum_j ext[j][rows][N] * evk[j][2][evk_rows][N]; ERow maps working row -> evk row.  The key is read
// ONCE per coefficient and applied to every batch item (its HBM traffic is amortised over the batch).
void launch_ks_inner(KShape S, u64* acc, const u64* ext, const u64* own, const u64* evk, const u32* perm, const LimbList& L, const LimbList& ERow, int beta, int evk_rows, int nq, int alpha, const u64* addend, const ScalarList& PmodQ, int accumulate, dev_stream st, bool tensor = false, KsBatch kb = KsBatch());
// nz independent conversions: slice z reads in + z*in_zs, writes out + z*out_zs, with table tabs_dev[z*tab_zstride]
// (device memory); every table of the launch has exactly `ns` sources and at most `max_nt` targets.  nb batch items
// repeat the nz slices at in + b*in_bs / out + b*out_bs.
void launch_base_convert(KShape S, u64* out, const u64* in, const BaseConvTable* tabs_dev, int tab_zstride, int ns, int max_nt, int nz, size_t in_zs, size_t out_zs, dev_stream st, int mode = BC_INT, int nb = 1, size_t in_bs = 0, size_t out_bs = 0);
void launch_rescale_delta(KShape S, u64* delta, const u64* last, const LimbList& L, int last_mod, int npoly, PolyStride ps, dev_stream st);
void launch_center_lift(KShape S, u64* out, const u64* in, const LimbList& L, int src_mod, int npoly, PolyStride ps, dev_stream st);
void launch_sample_uniform(KShape S, u64* out, const LimbList& L, u64 seed, u64 stream, dev_stream st);
// nb batch items: item b writes out + b*out_bs and draws from stream + (b << 16) (the "a" field of the stream id).  epoch
// (device word, may be null): the seed is offset by *epoch * const, so that a captured graph draws fresh randomness on
// every replay (the engine bumps the word at the head of each captured graph; 0 outside graphs: oracle parity).
void launch_sample_small(KShape S, u64* out, const LimbList& L, u64 seed, u64 stream, int kind, dev_stream st, int nb = 1, size_t out_bs = 0, const u64* epoch = nullptr);
void launch_bump(u64* word, dev_stream st);
void launch_reduce_i64(KShape S, u64* out, const i64* v, const LimbList& L, dev_stream st, int nb = 1, size_t out_bs = 0);
// embedding kernels: nb independent vectors, consecutive in memory (n complex slots / N coefficients each)
void launch_special_fft(KShape S, double* out, const double* w, const u32* rot, const double* ksi, dev_stream st, int nb = 1);
void launch_special_ifft(KShape S, double* out, double* z, const u32* rot, const double* ksi, dev_stream st, int nb = 1);
void launch_snap_zeta16(KShape S, double* z, const double* table, int stride, dev_stream st, int nb = 1);
void launch_round_coeffs(KShape S, i64* out, const double* w, double scale, int* flag, dev_stream st, int nb = 1);
void launch_center_to_w(KShape S, double* w, const u64* coef, int mod, double scale, dev_stream st, int nb = 1);
void launch_zeta16_from_nibbles(KShape S, double* z, const unsigned char* nib, const double* table, dev_stream st, int nb = 1);
void launch_nibbles_from_zeta16(KShape S, unsigned char* nib, const double* z, dev_stream st, int nb = 1);
