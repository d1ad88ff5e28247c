// ntt.cuh -- host entry points of the negacyclic NTT (see ntt.cu).
#pragma once
#include "common.cuh"

// device tables, [n_moduli][N] each: psi powers in bit-reversed order and their Shoup companions
struct NttTables {
    const u64* fwd;
    const u64* fwd_s;
    const u64* inv;
    const u64* inv_s;
    // FP64 path (moduli below CKKS_FP_LIMIT): the same twiddles as doubles (no companion: common.cuh modmul_fp)
    const double* fwd_d;
    const double* inv_d;
    const ModConst* mc;
    int logn;
    int cluster;     // 0: two-pass kernels; 1: single-kernel (8-CTA cluster) forward transform at N = 2^16; 2: also the fused
                     // forward variants and the inverse.  Per engine (CKKS_NTT_CLUSTER is read by the Engine constructor).
};

// One batched transform: for z-slice z, item i reads source row srows[z][i] of src + z*szs and
// writes destination row rows[z][i] of dst + z*dzs (elements); the residues are modulo
// mc[mods[z][i]].  The first pass goes src -> dst, the second runs in place on dst; src may equal dst
// when srows == rows.
#define NTT_MAX_Z 6
// Batched ciphertexts: the whole job is repeated for nb batch items at src + b*sbs / dst + b*dbs (the grid's z extent is
// nz * nb, batch-major); nb = 0 means 1.
struct NttJob {
    int n, nz;
    int nb;
    size_t szs, dzs;
    size_t sbs, dbs, s2bs;            // batch strides (elements) of src, dst and the second source of ntt_inverse
    unsigned char mods[NTT_MAX_Z][CKKS_MAX_MODULI];
    unsigned char rows[NTT_MAX_Z][CKKS_MAX_MODULI];
    unsigned char srows[NTT_MAX_Z][CKKS_MAX_MODULI];
    unsigned char cnt[NTT_MAX_Z];     // items of slice z (0 = all n); CTAs beyond it exit at once
};

// Work fused into the forward transform (one launch less each, and the intermediate never goes to HBM):
//   prologue (pass A, pro_mod >= 0): every source row holds the SAME kind of data, one coefficient-domain limb modulo
//     q_l = mc[pro_mod]; item i transforms the centred rescale lift ((x + h) mod q_l) mod q_i - (h mod q_i), h = q_l >> 1
//     (spec S6; the stand-alone kernel is k_rescale_delta);
//   epilogue (pass B, ep_out != null): instead of storing NTT(x), item i of slice z stores (a - NTT(x)) * s[i] into row
//     rows[z][i] of ep_out + z*ep_ozs, a = the same row of ep_a + z*ep_azs (the tail of ModDown and of a rescale; the
//     stand-alone kernel is k_sub_mul_scalar).
//     With ep_k the scalar multiplication of a level alignment rides along: (ep_k * a - NTT(x)) * s[i].
struct NttFuse {
    int pro_mod;
    const u64* ep_a;
    u64* ep_out;
    size_t ep_azs, ep_ozs;
    size_t ep_abs = 0, ep_obs = 0;    // batch strides of ep_a / ep_out
    ScalarList s;
    u64 ep_k = 0;                     // != 0: a is multiplied by this integer first, (k a - NTT(x)) * s (level alignment, spec S6)
    ScalarList kl;                    // ... as residues of item i's modulus with their Shoup companions (filled when ep_k != 0)
};

void ntt_forward(const u64* src, u64* dst, const NttJob& J, const NttTables& T, dev_stream st);
void ntt_forward_fused(const u64* src, u64* dst, const NttJob& J, const NttTables& T, const NttFuse& F, dev_stream st);
// src2 != null: the inverse transform of the row-wise product src * src2 (the d2 of a ct x ct, never stored)
void ntt_inverse(const u64* src, u64* dst, const NttJob& J, const NttTables& T, dev_stream st, const u64* src2 = nullptr);
