// engine.cuh -- host-side CKKS evaluation engine driving the sm_100a kernels.
//
// One Engine per GPU, one CUDA stream per Engine; every operation is enqueued asynchronously
// and is out-of-place (inputs are immutable: the reference callers alias ciphertexts freely,
// SURVEY.md 8b).  Polynomials live in HBM as [limb][N] uint64 residues in the NTT domain
// (bit-reversed order); a ciphertext is [npoly][level+1][N] contiguous.  Scratch and results
// come from the stream-ordered CUDA memory pool (no cudaMalloc on the hot path).
//
// Scale policy (DESIGN.md spec S1): one canonical scale per level, S[L] = 2^scale_bits,
// S[l-1] = S[l]^2 / q_l.  Every ciphertext at level l has exactly scale S[l]; plaintext operands
// of a multiplication are encoded at S[l] so the rescaled product lands on S[l-1]; operands of
// different levels are aligned by level_down() (limb drop + one integer multiply + one rescale).
//
// Batch dimension.  A Ct may hold nb INDEPENDENT ciphertexts of one shape, [nb][npoly][level+1][N] contiguous (the AES
// path: one item per packed ciphertext pair, BASELINE.json configs[4]).  Every operation acts on all items in ONE set of
// kernel launches -- an NTT call carries nb x the limbs, the key of a key switch is read once for the whole batch -- and
// item i of the result is bit-identical to the operation applied to item i alone.  An nb = 1 operand of a binary
// operation is broadcast to the batch (round-key ciphertexts are shared by all pairs).
#pragma once
#include <map>
#include <unordered_map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "ntt.cuh"

namespace ckks {

// error classes the C ABI maps to status codes (and the Python shim to RuntimeError strings,
// reference engine_context.py:139-145,184-195)
struct LevelError : std::runtime_error { using std::runtime_error::runtime_error; };
struct FormError : std::runtime_error { using std::runtime_error::runtime_error; };
struct PolyCountError : std::runtime_error { using std::runtime_error::runtime_error; };

struct Ct {
    int npoly = 2;
    int level = 0;
    int nb = 1;                                     // batch items (independent ciphertexts of this shape)
    u64* d = nullptr;                               // [nb][npoly][level+1][N], NTT domain
    int lane = 0;                                   // stream lane that produced it (0: the main stream)
    long epoch = 0;                                 // fork region it was produced in
    int cap = 0;                                    // graph capture it was recorded in (0: produced eagerly)
    std::vector<std::pair<int, Ct*>> lowered;       // memoised level_down() results (owned)
};
struct Pt {
    int level = 0;
    u64* d = nullptr;                               // [level+1][N], NTT domain, scale S[level]
};
struct EvalKey {
    u64* d = nullptr;                               // [dnum][2][L+1+K][N]
};
// ModUp result: every digit of a polynomial extended to Q_level u P, NTT domain
struct Decomp {
    int level = 0, beta = 0;
    int nb = 1;
    u64* ext = nullptr;                             // [nb][beta][level+1+K][N]; a digit's own rows are not filled
    const u64* own = nullptr;                       // the NTT-domain input [level+1][N] (read for own rows)
    size_t own_bs = 0;                              // element stride of `own` between batch items (0: shared)
};

struct BootParams {
    int enabled = 0;
    int cts_groups = 3, stc_groups = 3;             // matrices the (i)DFT is factored into
    int K = 25;                                     // bound on |I| in t = m + q0 I
    int cheb_degree = 47;                           // 63 gives the same measured precision (5.6e-4) for 4 more products
    int double_angle = 3;
};

struct Params {
    int logn = 16;
    std::vector<u64> q, p;                          // q_0..q_L ; p_0..p_{K-1}
    int scale_bits = 50;
    int top_levels = 0, top_bits = 0;               // S[L] = 2^top_bits when top_levels > 0 (see default_params)
    double top_scale = 0;                           // S[L] as default_params chose it (0: 2^top_bits or 2^scale_bits)
    int scale_drop = 0;                             // bits S_0 sits below 2^scale_bits (descending-scale chain, q_0 < 2^(scale_bits+10))
    int alpha = 1;                                  // q-limbs per key-switch digit
    int hamming = 192;
    int fresh_level = 0;                            // level of fresh encryptions (<= L)
    u64 seed = 1;
    int device = 0;
    BootParams boot;
};

// Deterministic prime chain (DESIGN.md spec S1; restated independently in oracle/params.py)
Params default_params(int logn, int levels, int scale_bits, int q0_bits, int p_bits, int dnum, int hamming,
                      int fresh_level, int top_levels = 0, int top_bits = 0);

struct BootPlan;   // bootstrap.cu

class Engine {
  public:
    explicit Engine(const Params& P);
    ~Engine();
    Engine(const Engine&) = delete;

    // ---- parameters
    int L() const { return (int)prm.q.size() - 1; }
    int K() const { return (int)prm.p.size(); }
    int nmod() const { return L() + 1 + K(); }
    int dnum() const { return (L() + 1 + prm.alpha - 1) / prm.alpha; }
    size_t N() const { return (size_t)1 << prm.logn; }
    size_t slots() const { return N() / 2; }
    double scale_at(int level) const { return scales[level]; }
    Params prm;
    std::vector<double> scales;
    dev_stream st = 0;

    // ---- memory
    u64* alloc(size_t words);
    void release(void* p);
    Ct* new_ct(int npoly, int level, int nb = 1);
    static size_t bstride(const Ct* c, size_t per_item) { return c->nb > 1 ? per_item : 0; }   // 0 broadcasts an nb = 1 operand
    int batch_of(const Ct* a, const Ct* b) const;   // common batch size of a binary operation (nb = 1 broadcasts)
    Ct* stack(const std::vector<Ct*>& items);       // ciphertexts of one shape (batched ones contribute all their items) -> one batched ciphertext (copies)
    Ct* slice(const Ct* c, int start, int count);   // copy of batch items start .. start + count - 1
    Ct* item(const Ct* c, int i);                   // copy of batch item i as an nb = 1 ciphertext
    void free_ct(Ct* c);
    void free_pt(Pt* p);
    void sync();

    // ---- lanes: independent pieces of work (the hi / lo nibble planes, the two power bases of an XOR4, the products of
    // one generation of a power basis) are enqueued on separate CUDA streams so the device overlaps their small kernels.
    // fork(k) takes k helper streams ordered after the current one, set_lane(i) routes the following calls to lane i,
    // join() orders the parent after all lanes.  Forks nest.  A ciphertext released while another lane may still read it
    // is held until the join of the frame.
    static const int kMaxLanes = 40;
    struct LanePool { std::map<size_t, std::vector<void*>> free; size_t cached = 0; };
    struct Frame {
        int parent = 0;
        long epoch = 0;
        bool serial = false;
        std::vector<int> lanes;
        std::vector<Ct*> dct;
        std::vector<Pt*> dpt;
    };
    void fork(int k = 2);
    void set_lane(int i);
    void join();
    bool in_fork() const { return !frames.empty(); }
    bool lanes_on = true;                  // false: forks run serially on the parent stream (A/B timing, profiling)
    dev_stream streams[kMaxLanes] = {};
    bool lane_made[kMaxLanes] = {};
    bool lane_busy[kMaxLanes] = {};
    int cur_lane = 0;                      // 0 = the main stream
    long epoch_counter = 0;
    std::vector<Frame> frames;
    long cur_epoch() const { return frames.empty() ? 0 : frames.back().epoch; }
    // stream-ordered caching arena: a released buffer goes to the free list of the lane it is released on and is handed
    // out again to later work on the SAME stream (stream order makes that safe), so the hot path makes no
    // cudaMallocAsync / cudaFreeAsync calls at all; fork deals the parent's cached buffers to the lanes, join returns them
    struct MemArena {
        int id = 0;                        // 0: the engine's own arena; otherwise the graph that owns it
        LanePool pools[kMaxLanes];
        std::vector<void*> plain;          // cudaMalloc'ed during a capture (an arena miss must not become a graph node)
        long capture_misses = 0;
        size_t bytes = 0;
    };
    struct Buf { size_t bytes; int arena; };
    MemArena main_arena;
    MemArena* arena = &main_arena;
    std::map<int, MemArena*> arenas;          // live arenas by id (the main arena is looked up by pointer)
    std::map<void*, Buf> alloc_bytes;
    void trim_pools();

    // ---- captured graphs (CUDA graphs): a whole sequence of engine calls -- e.g. one AES round, 13 000 launches over
    // nested stream lanes -- is recorded once into a private arena and replayed with one driver call; several graphs
    // replay concurrently on separate replay streams (one ciphertext pair each)
    static const int kMaxReplay = 16;
    struct GraphRec {
        int id = 0;
        MemArena arena;
        dev::Graph* g = nullptr;
        long launches0 = 0, launches = 0, replays = 0;
        long cnt0[5] = {}, cnt[5] = {};
    };
    std::map<int, GraphRec*> graphs;
    int graph_counter = 0;
    int capture_id = 0;                    // graph whose capture is open (0: none)
    dev_stream replay[kMaxReplay] = {};
    int graph_create();
    GraphRec* graph_rec(int id);
    void graph_enter(int id);
    void graph_leave();
    void graph_capture_begin(int id);
    void graph_capture_end(int id);
    void graph_capture_abort();
    void graph_launch(int id, int slot);
    void graph_wait(int slot);
    void graph_destroy(int id);
    dev_stream replay_stream(int slot);
    void ct_assign(Ct* dst, const Ct* src, int slot);
    void ct_clear_memo(Ct* c);
    long n_driver_allocs = 0;
    size_t driver_bytes = 0;

    // ---- keys (spec S8)
    void keygen_secret();
    void keygen_public();
    void keygen_relin();
    EvalKey* galois_key(u64 g);                 // generated on first use from sk, cached
    u64 galois_for_rotation(long steps) const;  // rotate(ct,+r) == np.roll(slots,+r)
    u64 galois_conj() const { return 2ull * N() - 1; }
    bool has_sk = false, has_pk = false, has_relin = false;
    bool keys_external = false;                 // switching keys are allocated but not sampled: filled by a broadcast
    std::vector<u64> switch_key_ids() const;    // 0 = relinearisation key, else Galois element
    u64* switch_key_buffer(u64 id, size_t* words);

    // ---- encode / encrypt / decrypt (host pointers: interleaved re,im doubles, n = N/2 slots)
    // batched forms: z / nib hold nb consecutive slot vectors; decrypt writes c->nb of them
    Pt* encode(const double* z, int level);
    Ct* encrypt(const double* z, int level, int nb = 1);
    void decrypt(const Ct* c, double* z_out);
    double* decrypt_to_dev(const Ct* c);
    Ct* encrypt_coeffs(const i64* coef_dev, int level, int nb = 1);
    Ct* encrypt_zeta16(const unsigned char* nib_host, int level, int nb = 1);   // one nibble per slot in, zeta_16 codewords encrypted
    void decrypt_zeta16(const Ct* c, unsigned char* nib_out_host);  // nearest zeta_16 codeword index per slot out
    Ct* snap_zeta16(const Ct* a, int level, int stride);   // decrypt, snap to zeta_16 codewords, re-encrypt: all on device

    // ---- homomorphic ops (all out of place)
    Ct* add(Ct* a, Ct* b);
    Ct* sub(Ct* a, Ct* b);
    Ct* negate(const Ct* a);
    Ct* mul(Ct* a, Ct* b, int factor = 1);                       // tensor + relinearise + rescale
    Ct* mul_norelin(Ct* a, Ct* b);               // tensor + rescale, 3 polynomials
    Ct* relinearize(const Ct* t);                // 3 -> 2 polynomials, same level
    Ct* rescale(const Ct* c);
    Ct* level_down(Ct* c, int target);           // memoised on c
    Ct* lowered_copy(const Ct* c, int target);          // the same alignment, not memoised (caller owns it)
    Ct* mul_const(const Ct* a, double re, double im);   // consumes one level
    Ct* mul_plain(const Ct* a, const Pt* p);            // consumes one level; p at a.level
    Ct* mul_i(const Ct* a, int sign);                   // exact multiply by +-i, no level
    Ct* add_const(const Ct* a, double re, double im);
    Ct* add_plain(const Ct* a, const Pt* p);
    Ct* rotate(const Ct* a, long steps);
    Ct* conjugate(const Ct* a);
    Ct* apply_galois(const Ct* a, u64 g);
    std::vector<Ct*> rotate_hoisted(const Ct* a, const std::vector<long>& steps);
    std::vector<Ct*> power_basis(Ct* a, int degree, const unsigned char* need = nullptr);
    Ct* copy(const Ct* a);
    Ct* lut2(const std::vector<Ct*>& A, const std::vector<Ct*>& B, const int* p, const int* q, const double* coef,
             int nterms);
    Ct* lincomb(const std::vector<Ct*>& X, const double* coef, int n);
    void diag_mac(u64* out, const std::vector<const Ct*>& x, const std::vector<const Pt*>& p, int level, int nb = 1);
    void diag_mac_rows(const std::vector<u64*>& out, const std::vector<const Ct*>& x,
                       const std::vector<std::vector<const Pt*>>& p, int level, int nb = 1);
    const u64* const_table(const double* coef_re_im, int n, int scale_level, int level);
    Ct* drop_to(const Ct* a, int level);         // plain limb drop (scale unchanged): internal/bootstrap use

    // ---- bootstrapping (bootstrap.cu)
    void bootstrap_setup();
    void bootstrap_teardown();
    Ct* bootstrap(Ct* a);
    int boot_out_level() const;
    Ct* mod_raise(Ct* a);
    std::unique_ptr<BootPlan> boot;

    // ---- low-level pieces (exposed through the C ABI for parity tests against the oracle)
    void ntt_rows(u64* data, const std::vector<int>& rows, const std::vector<int>& mods, bool inverse, int nz = 1,
                  size_t zstride = 0, int nb = 1, size_t bstride = 0);
    void run_ntt(const u64* src, u64* dst, const NttJob& J, bool inverse, long limbs, const u64* src2 = nullptr);
    void run_ntt_fused(const u64* src, u64* dst, const NttJob& J, const NttFuse& F, long limbs);
    // basis conversion: BC_FP (default) = scale-prime targets on the FP64 pipe, 60/61-bit targets on the integer pipe;
    // CKKS_BC_FP=0: every target on the integer pipe; CKKS_BC_MMA=1: byte-sliced u8 tensor-core GEMM (measured slower: profiles/README.md)
    int bc_mode = BC_FP;
    int bc_fp_int_every = 0;               // CKKS_BC_FP_INT_EVERY=k: every k-th FP64-capable target stays on the integer pipe
    bool fuse_tensor = true;               // CKKS_TENSOR_FUSE=0: ct x ct writes its tensor product (k_tensor) first
    bool fuse_ntt = true;                  // CKKS_NTT_FUSE=0: stand-alone lift / subtract-scale kernels (A/B timing)
    void profile_begin();
    void profile_end(double* ms, long* calls, long* limbs);
    // batched forms: nb items, d / times / addend advance by their batch stride per item (0 = shared by all items);
    // ext, acc and out are [nb][...] contiguous
    Decomp decompose(const u64* d, int level, const u64* times = nullptr, int nb = 1, size_t d_bs = 0, size_t times_bs = 0);
    void ks_apply(const Decomp& D, const EvalKey* evk, const u32* perm, u64* out /* [nb][2][level+1-drop][N] */,
                  const u64* addend = nullptr, int drop = 0, bool tensor = false, size_t addend_bs = 0, int addend_mode = 0,
                  int factor = 1);
    // c0_addend (optional): a polynomial [level+1][N] per item (stride c0_bs) that joins output polynomial 0 (KsBatch::addend_mode 1)
    void key_switch(const u64* d, int level, const EvalKey* evk, u64* out, int nb = 1, size_t d_bs = 0,
                    const u64* c0_addend = nullptr, size_t c0_bs = 0);
    void ks_inner(const Decomp& D, const EvalKey* evk, const u32* perm, u64* acc, const u64* addend, bool accumulate,
                  bool tensor = false, size_t addend_bs = 0, int addend_mode = 0);
    bool fuse_align = true;                // CKKS_ALIGN_FUSE=0: level alignment as k_mul_scalar on every limb, then a rescale (A/B)
    bool fuse_mul_factor = true;           // CKKS_MUL_FACTOR_FUSE=0: 2 a b of the Chebyshev recurrences as a product and an addition (A/B)
    bool fuse_ks_add = true;               // CKKS_KS_ADD_FUSE=0: sigma(c0) + ks0 as a separate permutation / addition (A/B)
    void ks_moddown(u64* acc, int level, int drop, u64* out, int nb = 1, int factor = 1);
    void automorph(u64* out, const u64* in, int rows, int npoly, u64 g);
    void automorph(u64* out, const u64* in, int rows, int npoly, u64 g, PolyStride ps);
    const u32* galois_perm(u64 g);
    std::vector<int> mods_q(int level) const;
    std::vector<int> mods_qp(int level) const;
    LimbList limb_list(const std::vector<int>& mods) const;
    void const_residues(double re, double im, double scale, const std::vector<int>& mods, ScalarList& cp,
                        ScalarList& cm) const;
    void scalar_list(const std::vector<u64>& vals, const std::vector<int>& mods, ScalarList& out) const;
    const BaseConvTable& modup_table(int level, int digit);
    const BaseConvTable& moddown_table(int level, int drop = 0);
    const ScalarList& moddown_inv(int level, int drop, int factor = 1);
    const BaseConvTable* modup_tables_dev(int level);
    const BaseConvTable* moddown_table_dev(int level, int drop = 0);
    const std::vector<i64>& sk_host() const { return sk_coef; }
    const u64* sk_dev() const { return sk_ntt; }
    const u64* pk_dev() const { return pk; }
    const EvalKey* relin_key() const { return &relin; }
    const std::vector<u64>& moduli() const { return mod; }
    NttTables tabs{};
    KShape ks{};
    long n_keyswitch = 0, n_ntt_limbs = 0, n_rescale = 0, n_mul_cc = 0, n_boot = 0;

  private:
    std::vector<u64> mod;                  // all moduli, global index
    std::vector<u64> psi;
    std::vector<u64> Jroot;                // psi^(N/2)
    std::vector<i64> sk_coef;
    u64 *d_fwd = nullptr, *d_fwd_s = nullptr, *d_inv = nullptr, *d_inv_s = nullptr;
    ModConst* d_mc = nullptr;
    u64* sk_ntt = nullptr;                 // [nmod][N]
    u64* pk = nullptr;                     // [2][L+1][N]
    EvalKey relin;
    std::map<u64, EvalKey> gkeys;
    std::map<u64, u32*> perms;
    std::map<std::pair<int, int>, BaseConvTable> modup_tabs;
    std::map<int, BaseConvTable> moddown_tabs;
    std::map<int, BaseConvTable*> modup_dev, moddown_dev;   // device copies ([beta] per level / one per level)
    std::map<u64, u64*> const_tabs;                 // LUT constants by content hash
    std::map<u64, unsigned char*> index_tabs;       // LUT term index lists
    std::vector<void*> owned;              // device tables freed in the destructor
    u32* d_rot = nullptr;
    double* d_ksi = nullptr;
    double* d_zeta16 = nullptr;
    int* d_flag = nullptr;
    u64 enc_counter = 0;
    u64* d_epoch = nullptr;                // replay epoch of captured graphs (mixed into the encryption randomness)
    bool prof_on = false;
    std::vector<dev::Timer> prof_timers;
    size_t prof_used = 0;
    long prof_limbs = 0, prof_calls = 0;
    double prof_ms = 0;
    ScalarList sl_pinv;                    // P^-1 mod q_i
    ScalarList sl_pmodq;                   // P mod q_i
    std::map<int, ScalarList> moddown_invs; // (level, drop, factor) -> factor (P q_dropped)^-1 mod q_i
    std::vector<ScalarList> sl_qinv;       // [l]: q_l^-1 mod q_i, i < l

    EvalKey make_switch_key(u64 key_id, const u64* s_from_ntt);
    BaseConvTable make_bc_table(const std::vector<int>& src, const std::vector<int>& srow,
                                const std::vector<int>& tgt, const std::vector<int>& orow, bool exact = false);
    void encode_coeffs_dev(i64* out_dev, const double* z_host, double scale, int nb = 1);
    void encode_coeffs_from_dev(i64* out_dev, double* z_dev, double scale, bool check, int nb = 1);
    void rescale_into(u64* out, const u64* in, int npoly, int level, int nb = 1, size_t in_ps = 0, u64 pre_k = 0);
    void need_levels(int level, int need, const char* what) const;
};

u64 mulmod_h(u64 a, u64 b, u64 q);
u64 powmod_h(u64 a, u64 e, u64 q);
u64 invmod_h(u64 a, u64 q);
u64 shoup_h(u64 w, u64 q);
u64 rand64_h(u64 seed, u64 stream, u64 idx);


}  // namespace ckks
