/* ckks_b200.h -- C ABI of the B200-native CKKS evaluation engine (libckks_b200.so).
 *
 * Drop-in boundary for the one thing the reference's AES-on-CKKS stack calls: the `desilofhe.Engine`
 * methods reached through reference `engine_context.py` (the only file that touches the backend,
 * SURVEY.md 8b).  Each entry point cites the reference interface it replaces.  Plain pointers,
 * sizes and opaque handles only -- no torch / C++ types.  The Python shim
 * `aes-implementation-fhe_b200/desilofhe/` binds these with ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - Every function returns an int status: 0 ok, CKKS_ERR_LEVEL (levels exhausted: the shim raises
 *     RuntimeError("... level should be positive ..."), the string the reference's recovery ladders
 *     match, xor4_lut.py:33-51, engine_context.py:184-195), CKKS_ERR_FORM ("... NTT ..."),
 *     CKKS_ERR_POLYS ("... should have 3 polynomials", engine_context.py:139-145), CKKS_ERR_OTHER.
 *     ckks_last_error() returns the message of the last failure on the calling thread.
 *   - All operations are out of place; inputs are never modified (the reference aliases
 *     ciphertexts freely).  Results are new handles owned by the caller (ckks_ct_free / ckks_pt_free).
 *   - Slot vectors are interleaved (re, im) doubles, slot_count = N/2 complex values.
 *   - One engine = one GPU + one CUDA stream; calls are asynchronous on that stream except
 *     ckks_decrypt / ckks_ct_export / ckks_sync, which synchronise.
 */
#ifndef CKKS_B200_H
#define CKKS_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ckks_engine ckks_engine;
typedef struct ckks_ct ckks_ct;     /* ciphertext: 2 (or 3, before relinearisation) polynomials over Q_level */
typedef struct ckks_pt ckks_pt;     /* encoded plaintext polynomial at one level */

enum { CKKS_OK = 0, CKKS_ERR_LEVEL = 1, CKKS_ERR_FORM = 2, CKKS_ERR_POLYS = 3, CKKS_ERR_OTHER = 4 };

const char* ckks_last_error(void);
const char* ckks_backend(void);          /* "cuda-sm_100a" for the product library */
long ckks_launch_count(void);            /* kernels launched by this library so far (bench.py gpu_launches) */

/* ---- engine construction: replaces desilofhe.Engine(...) (engine_context.py:17-42).
 * `_default` derives the deterministic prime chain of DESIGN.md spec S1 (q0 | L scale primes | K special); with
 * top_levels > 0 the highest levels carry the larger scale 2^top_bits (bootstrapping precision). */
int ckks_engine_create_default(int logn, int levels, int scale_bits, int q0_bits, int p_bits, int dnum,
                               int hamming_weight, int fresh_level, int top_levels, int top_bits, uint64_t seed,
                               int device_id, ckks_engine** out);
int ckks_engine_create(int logn, const uint64_t* q, int nq, const uint64_t* p, int np, int scale_bits, int alpha,
                       int hamming_weight, int fresh_level, uint64_t seed, int device_id, ckks_engine** out);
void ckks_engine_destroy(ckks_engine* e);
int ckks_sync(ckks_engine* e);
/* Stream lanes for independent pieces of work (the hi- and lo-nibble ciphertexts never interact inside XOR4 / bootstrap,
 * xor4_lut.py:63-74, mixcol_final.py:158-163; the two power bases of an XOR4 are independent, xor4_lut.py:65-66):
 * fork(k) orders k helper streams after the current one, set_lane(i) routes the following calls to lane i, join orders
 * the parent stream after all lanes.  Forks nest. */
int ckks_fork(ckks_engine* e, int lanes);
int ckks_set_lane(ckks_engine* e, int lane);
int ckks_set_lanes_enabled(ckks_engine* e, int on);   /* 0: forks run serially on the parent stream (profiling, A/B) */
int ckks_join(ckks_engine* e);
/* engine.slot_count (read at pipeline.py:39, xor4_lut.py:16, state_encoder.py:14, ...) */
int ckks_slot_count(const ckks_engine* e);
/* parameter introspection: any out pointer may be NULL; q_out/p_out/scales_out need nq/np/nq entries */
int ckks_get_params(const ckks_engine* e, int* logn, int* nq, int* np, int* alpha, int* fresh_level,
                    uint64_t* q_out, uint64_t* p_out, double* scales_out);

/* ---- keys: create_secret_key / create_public_key / create_relinearization_key /
 * create_conjugation_key / create_rotation_key / create_bootstrap_key (engine_context.py:44-50).
 * Keys live inside the engine; Galois keys for other steps are derived on first use. */
int ckks_keygen_secret(ckks_engine* e);
int ckks_keygen_public(ckks_engine* e);
int ckks_keygen_relin(ckks_engine* e);
int ckks_keygen_conjugation(ckks_engine* e);
int ckks_keygen_rotation(ckks_engine* e, const long* steps, int nsteps);
int ckks_keygen_bootstrap(ckks_engine* e);
/* Multi-GPU key distribution (one broadcast of the evaluation keys from rank 0 at setup, no per-round
 * exchange).  With external = 1 the keygen_* calls for switching keys (relin, conjugation, rotation, bootstrap) only
 * allocate; the caller fills the buffers (NCCL broadcast over NVLink) before the first evaluation.  The secret and
 * public keys are always derived from the seed. */
int ckks_set_keys_external(ckks_engine* e, int external);
int ckks_switch_key_ids(ckks_engine* e, uint64_t* ids_out, int capacity, int* count);   /* 0 = relinearisation key */
int ckks_switch_key_buffer(ckks_engine* e, uint64_t id, void** device_ptr, size_t* bytes);
/* optional, before ckks_keygen_bootstrap: |I| bound K, Chebyshev degree, double-angle steps, matrices per DFT */
int ckks_set_bootstrap_params(ckks_engine* e, int K, int cheb_degree, int double_angle, int cts_groups, int stc_groups);

/* ---- data movement: encode / encrypt / decrypt (engine_context.py:56-63) */
int ckks_encode(ckks_engine* e, const double* slots_re_im, int level, ckks_pt** out);
int ckks_encrypt(ckks_engine* e, const double* slots_re_im, int level /* <0: fresh level */, ckks_ct** out);
int ckks_decrypt(ckks_engine* e, const ckks_ct* ct, double* slots_re_im_out);
/* The reference's hard renorm (pipeline.py:65-69: decrypt, snap every nibble to its zeta_16 codeword, re-encrypt) done
 * without leaving the device: no D2H/H2D of the 512 KiB slot vector.  stride > 1 reproduces StateEncoder's layout
 * (data on slots 0 mod stride, 1.0 elsewhere, state_encoder.py:23-27); level < 0 means the fresh level. */
int ckks_snap_zeta16(ckks_engine* e, const ckks_ct* ct, int level, int stride, ckks_ct** out);
/* zeta_16 codec next to the data (the host loops of state_encoder.py:14-38 / utils.py:9-19): `nibbles` holds one value
 * 0..15 per slot (slot_count bytes); encryption looks the codewords exp(-2 pi i k / 16) up on the device, decryption
 * returns the index of the nearest codeword of every slot.  16 x fewer bytes over PCIe than complex128 slots. */
int ckks_encrypt_zeta16(ckks_engine* e, const uint8_t* nibbles, int level, ckks_ct** out);
int ckks_decrypt_zeta16(ckks_engine* e, const ckks_ct* ct, uint8_t* nibbles_out);
/* ---- batched ciphertexts (BASELINE.json configs[4]: "many ciphertexts"; reference test/test_aes_pipeline_roundtrip.py
 * processes ONE ciphertext pair).  A handle may hold nb INDEPENDENT ciphertexts of one shape; every operation below
 * (arithmetic, key switches, LUTs, bootstrap, snap) then acts on all items with one set of kernel launches, item i of the
 * result being bit-identical to the operation on item i alone; an nb = 1 operand of a binary operation is broadcast
 * (a round-key ciphertext shared by all pairs).  The *_batch encryptions take nb consecutive slot / nibble vectors;
 * ckks_decrypt, ckks_decrypt_zeta16 and ckks_ct_export write ckks_ct_batch(ct) consecutive vectors. */
int ckks_ct_batch(const ckks_ct* ct);
int ckks_encrypt_batch(ckks_engine* e, const double* slots_re_im /* [nb][2 slot_count] */, int nb, int level, ckks_ct** out);
int ckks_encrypt_zeta16_batch(ckks_engine* e, const uint8_t* nibbles /* [nb][slot_count] */, int nb, int level, ckks_ct** out);
int ckks_ct_stack(ckks_engine* e, ckks_ct* const* items /* n ciphertexts of one shape; a batched one contributes all its items */, int n, ckks_ct** out);
int ckks_ct_item(ckks_engine* e, const ckks_ct* ct, int index, ckks_ct** out /* copy of one item, nb = 1 */);
/* copy of items start .. start + count - 1.  With ckks_ct_stack: the two nibble planes of an AES state (reference
 * mixcol_final.py:158-162 bootstraps them one after the other) go through ONE bootstrap of 2 nb items */
int ckks_ct_slice(ckks_engine* e, const ckks_ct* ct, int start, int count, ckks_ct** out);
void ckks_ct_free(ckks_engine* e, ckks_ct* ct);
void ckks_pt_free(ckks_engine* e, ckks_pt* pt);
int ckks_ct_level(const ckks_ct* ct);
int ckks_ct_npoly(const ckks_ct* ct);
int ckks_pt_level(const ckks_pt* pt);

/* ---- arithmetic: multiply / add / subtract / add_plain (engine_context.py:65-98).
 * Binary operations align levels themselves (xor4_lut.py:69-73 adds ciphertexts 5-13 levels apart);
 * every multiply consumes exactly one level and rescales. */
int ckks_add(ckks_engine* e, ckks_ct* a, ckks_ct* b, ckks_ct** out);
int ckks_sub(ckks_engine* e, ckks_ct* a, ckks_ct* b, ckks_ct** out);
int ckks_negate(ckks_engine* e, const ckks_ct* a, ckks_ct** out);
int ckks_mul(ckks_engine* e, ckks_ct* a, ckks_ct* b, ckks_ct** out);           /* multiply(a, b, relin_key) */
int ckks_mul_norelin(ckks_engine* e, ckks_ct* a, ckks_ct* b, ckks_ct** out);   /* multiply(a, b): 3 polynomials */
int ckks_relinearize(ckks_engine* e, const ckks_ct* a, ckks_ct** out);         /* engine_context.py:134-145 */
int ckks_mul_const(ckks_engine* e, const ckks_ct* a, double re, double im, ckks_ct** out);  /* multiply(ct, float | constant plaintext) */
int ckks_mul_plain(ckks_engine* e, const ckks_ct* a, const ckks_pt* p, ckks_ct** out);      /* multiply(ct, plaintext) */
int ckks_add_const(ckks_engine* e, const ckks_ct* a, double re, double im, ckks_ct** out);  /* add_plain(ct, float) */
int ckks_add_plain(ckks_engine* e, const ckks_ct* a, const ckks_pt* p, ckks_ct** out);      /* add(ct, plaintext) */
int ckks_mul_i(ckks_engine* e, const ckks_ct* a, int sign, ckks_ct** out);                  /* exact multiply by +-i */
int ckks_level_down(ckks_engine* e, ckks_ct* a, int level, ckks_ct** out);

/* make_power_basis(ct, degree, relin_key) -> [ct^1 .. ct^degree] (engine_context.py:100-101) */
int ckks_power_basis(ckks_engine* e, ckks_ct* a, int degree, ckks_ct** out /* degree handles */);
/* the same basis restricted to the listed exponents and the intermediates their products need (same products, so the
 * powers that are computed are bit-identical to ckks_power_basis'); handles of powers that were not computed are NULL.
 * The XOR4 table (xor4_lut.py:10-77) only has odd exponents: 5 products and 4 conjugations per base instead of 7 + 7. */
int ckks_power_basis_sparse(ckks_engine* e, ckks_ct* a, int degree, const int* exponents, int n, ckks_ct** out);
/* conjugate(ct, conj_key) (engine_context.py:103-104) */
int ckks_conjugate(ckks_engine* e, const ckks_ct* a, ckks_ct** out);
/* rotate(ct, rot_key, steps): out = np.roll(slots, steps) (engine_context.py:127-132, shift_rows.py:35-37) */
int ckks_rotate(ckks_engine* e, const ckks_ct* a, long steps, ckks_ct** out);
/* several rotations of one ciphertext sharing one ModUp (mixcol_final.py:124-126, invmixcolumns_fhe.py:140-142) */
int ckks_rotate_hoisted(ckks_engine* e, const ckks_ct* a, const long* steps, int nsteps, ckks_ct** out);

/* Fused sparse LUT evaluation (xor4_lut.py:63-74, mixcol_final.py:80-99, invmixcolumns_fhe.py:76-90):
 * out = sum_t c_t * A[p_t] * B[q_t] over the two 16-element power bases (NULL entries are unused),
 * one tensor accumulation, ONE relinearisation, two rescales: result at min level - 2. */
int ckks_lut2(ckks_engine* e, ckks_ct* const* A, ckks_ct* const* B, int nbasis, const int* p, const int* q,
              const double* coef_re_im, int nterms, ckks_ct** out);
/* Fused linear combination (sub_bytes_lut.py:49-54,63-71): out = sum_k c_k X[k]; one multiply-accumulate kernel and one
 * rescale per distinct input level instead of one multiply+rescale per term; result at (lowest input level - 1). */
int ckks_lincomb(ckks_engine* e, ckks_ct* const* X, int n, const double* coef_re_im /* [n][2] */, ckks_ct** out);

/* bootstrap(ct, relin, conj, bootstrap_key) (engine_context.py:147-162) */
int ckks_bootstrap(ckks_engine* e, ckks_ct* a, ckks_ct** out);
int ckks_bootstrap_out_level(const ckks_engine* e);

/* ---- counters (per engine): key switches, limb-NTTs, rescales, ct*ct multiplications, bootstraps */
int ckks_counters(const ckks_engine* e, long* out5);

/* arena statistics: driver allocations made so far, bytes held by the arena, bytes sitting in its free lists */
int ckks_arena_stats(const ckks_engine* e, long* driver_allocs, size_t* arena_bytes, size_t* cached_bytes);

/* ---- raw access for the bit-exact parity tests against oracle/ (tests/ only; not used by the shim's hot path) */
int ckks_ct_export(ckks_engine* e, const ckks_ct* ct, uint64_t* out /* [nb][npoly][level+1][N] */);
int ckks_ct_import(ckks_engine* e, int npoly, int level, const uint64_t* data, ckks_ct** out);
int ckks_ct_import_batch(ckks_engine* e, int nb, int npoly, int level, const uint64_t* data, ckks_ct** out);
int ckks_pt_export(ckks_engine* e, const ckks_pt* pt, uint64_t* out /* [level+1][N] */);
int ckks_export_secret(ckks_engine* e, int64_t* coef_out /* [N] */);
int ckks_export_public(ckks_engine* e, uint64_t* out /* [2][L+1][N] */);
int ckks_export_switch_key(ckks_engine* e, uint64_t galois_or_0_for_relin, uint64_t* out /* [dnum][2][L+1+K][N] */);
int ckks_test_ntt(ckks_engine* e, uint64_t* data /* [nrows][N], host, in place */, int nrows, const int* mods, int inverse);
int ckks_test_automorph(ckks_engine* e, uint64_t* data /* [nrows][N] host, in place */, int nrows, uint64_t galois);
int ckks_test_key_switch(ckks_engine* e, const uint64_t* poly /* [level+1][N] */, int level,
                         uint64_t galois_or_0_for_relin, uint64_t* out /* [2][level+1][N] */);
uint64_t ckks_galois_for_rotation(const ckks_engine* e, long steps);

/* ---- captured graphs (CUDA graphs).  The reference issues one backend call per homomorphic operation
 * (engine_context.py:65-162); one AES round is about 3 300 such calls and 13 000 kernel launches.  Because CKKS
 * evaluation is data-oblivious, the whole sequence can be recorded once and replayed with one driver call:
 *   create -> enter (a private arena becomes current) -> copy the inputs into static ciphertexts, run the calls once
 *   eagerly (fills the arena, builds every lazily created table), ckks_ct_clear_memo the static inputs ->
 *   capture_begin -> the same calls again (recorded, not executed; host<->device copies and synchronisation raise) ->
 *   capture_end -> leave.
 * ckks_ct_assign overwrites a static input, ckks_graph_launch replays on replay stream `replay_stream` (0 = the engine's
 * main stream; graphs on different replay streams run concurrently), ckks_graph_wait orders the main stream after it. */
int ckks_graph_create(ckks_engine* e, int* id_out);
int ckks_graph_enter(ckks_engine* e, int id);
int ckks_graph_leave(ckks_engine* e);
int ckks_graph_capture_begin(ckks_engine* e, int id);
int ckks_graph_capture_end(ckks_engine* e, int id);
int ckks_graph_capture_abort(ckks_engine* e);
int ckks_graph_launch(ckks_engine* e, int id, int replay_stream);
int ckks_graph_wait(ckks_engine* e, int replay_stream);
int ckks_graph_destroy(ckks_engine* e, int id);
int ckks_graph_info(ckks_engine* e, int id, long* nodes, long* launches, size_t* arena_bytes, long* capture_misses);
int ckks_ct_assign(ckks_engine* e, ckks_ct* dst, const ckks_ct* src, int replay_stream);
int ckks_ct_clear_memo(ckks_engine* e, ckks_ct* ct);

/* ---- timing helpers for bench.py: device-side timing on the engine's own stream (torch.cuda.Event only
 * sees torch's current stream) */
int ckks_timer_start(ckks_engine* e);
int ckks_timer_stop_ms(ckks_engine* e, float* ms_out);
/* per-call CUDA-event timing of every NTT launch between begin and end (bench.py roofline leg):
 * total milliseconds inside NTT kernels, number of batched calls, number of limb transforms */
int ckks_profile_ntt_begin(ckks_engine* e);
int ckks_profile_ntt_end(ckks_engine* e, double* ms_out, long* calls_out, long* limbs_out);
/* micro-benchmarks on resident random data: returns average milliseconds per call over `iters` calls */
int ckks_bench_ntt(ckks_engine* e, int nlimbs, int batches, int inverse, int iters, float* ms_out);
int ckks_bench_rotate(ckks_engine* e, int level, int iters, float* ms_out);
/* `lanes` independent ciphertexts rotated concurrently on stream lanes: milliseconds per rotation (throughput) */
int ckks_bench_rotate_lanes(ckks_engine* e, int level, int lanes, int iters, float* ms_per_rotation);
int ckks_bench_mul(ckks_engine* e, int level, int iters, float* ms_out);
/* one rotation / one ct*ct multiplication of a batched ciphertext of nb items: milliseconds per CALL (nb items each) */
int ckks_bench_rotate_batch(ckks_engine* e, int level, int nb, int iters, float* ms_out);
int ckks_bench_mul_batch(ckks_engine* e, int level, int nb, int iters, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif
