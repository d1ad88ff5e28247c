// capi.cu -- the C ABI of include/ckks_b200.h over ckks::Engine.
#include "../../include/ckks_b200.h"

#include "engine.cuh"

using namespace ckks;

struct ckks_engine {
    Engine* E;
    dev::Timer timer;
};
struct ckks_ct { Ct c; };
struct ckks_pt { Pt p; };

static thread_local std::string g_err;

template <typename F>
static int guard(F f) {
    try {
        f();
        return CKKS_OK;
    } catch (const LevelError& e) { g_err = e.what(); return CKKS_ERR_LEVEL;
    } catch (const FormError& e) { g_err = e.what(); return CKKS_ERR_FORM;
    } catch (const PolyCountError& e) { g_err = e.what(); return CKKS_ERR_POLYS;
    } catch (const std::exception& e) { g_err = e.what(); return CKKS_ERR_OTHER;
    } catch (...) { g_err = "unknown error"; return CKKS_ERR_OTHER; }
}
static inline Ct* C(ckks_ct* c) { return reinterpret_cast<Ct*>(c); }
static inline const Ct* C(const ckks_ct* c) { return reinterpret_cast<const Ct*>(c); }
static inline ckks_ct* H(Ct* c) { return reinterpret_cast<ckks_ct*>(c); }
static inline const Pt* P(const ckks_pt* p) { return reinterpret_cast<const Pt*>(p); }

extern "C" {

const char* ckks_last_error(void) { return g_err.c_str(); }
const char* ckks_backend(void) { return dev::backend_name(); }
long ckks_launch_count(void) { return g_launch_count; }
double ckks_launch_host_ms(void) {
#ifdef CKKS_TIME_LAUNCHES
    return g_launch_host_ns * 1e-6;
#else
    return -1.0;
#endif
}

int ckks_engine_create_default(int logn, int levels, int scale_bits, int q0_bits, int p_bits, int dnum, int hamming,
                               int fresh_level, int top_levels, int top_bits, uint64_t seed, int device,
                               ckks_engine** out) {
    return guard([&] {
        Params prm = default_params(logn, levels, scale_bits, q0_bits, p_bits, dnum, hamming, fresh_level, top_levels,
                                    top_bits);
        prm.seed = seed;
        prm.device = device;
        *out = new ckks_engine{new Engine(prm), {}};
    });
}
int ckks_engine_create(int logn, const uint64_t* q, int nq, const uint64_t* p, int np, int scale_bits, int alpha,
                       int hamming, int fresh_level, uint64_t seed, int device, ckks_engine** out) {
    // explicit chain: S[L] = 2^scale_bits, the lower scales follow from the primes (spec S1)
    return guard([&] {
        Params prm;
        prm.logn = logn;
        prm.q.assign(q, q + nq);
        prm.p.assign(p, p + np);
        prm.scale_bits = scale_bits;
        prm.alpha = alpha;
        prm.hamming = hamming;
        prm.fresh_level = fresh_level < 0 ? nq - 1 : fresh_level;
        prm.seed = seed;
        prm.device = device;
        *out = new ckks_engine{new Engine(prm), {}};
    });
}
void ckks_engine_destroy(ckks_engine* e) {
    if (!e) return;
    delete e->E;
    delete e;
}
int ckks_fork(ckks_engine* e, int lanes) { return guard([&] { e->E->fork(lanes); }); }
int ckks_set_lanes_enabled(ckks_engine* e, int on) {
    return guard([&] {
        if (e->E->in_fork()) throw std::runtime_error("cannot switch lanes inside a fork");
        e->E->lanes_on = on != 0;
    });
}
int ckks_set_lane(ckks_engine* e, int lane) { return guard([&] { e->E->set_lane(lane); }); }
int ckks_join(ckks_engine* e) { return guard([&] { e->E->join(); }); }
int ckks_sync(ckks_engine* e) { return guard([&] { e->E->sync(); }); }
int ckks_slot_count(const ckks_engine* e) { return (int)e->E->slots(); }
int ckks_get_params(const ckks_engine* e, int* logn, int* nq, int* np, int* alpha, int* fresh_level, uint64_t* q_out,
                    uint64_t* p_out, double* scales_out) {
    const Engine& E = *e->E;
    if (logn) *logn = E.prm.logn;
    if (nq) *nq = E.L() + 1;
    if (np) *np = E.K();
    if (alpha) *alpha = E.prm.alpha;
    if (fresh_level) *fresh_level = E.prm.fresh_level;
    if (q_out) for (int i = 0; i <= E.L(); i++) q_out[i] = E.prm.q[i];
    if (p_out) for (int i = 0; i < E.K(); i++) p_out[i] = E.prm.p[i];
    if (scales_out) for (int i = 0; i <= E.L(); i++) scales_out[i] = E.scales[i];
    return CKKS_OK;
}

int ckks_keygen_secret(ckks_engine* e) { return guard([&] { e->E->keygen_secret(); }); }
int ckks_keygen_public(ckks_engine* e) { return guard([&] { e->E->keygen_public(); }); }
int ckks_keygen_relin(ckks_engine* e) { return guard([&] { e->E->keygen_relin(); }); }
int ckks_keygen_conjugation(ckks_engine* e) { return guard([&] { e->E->galois_key(e->E->galois_conj()); }); }
int ckks_keygen_rotation(ckks_engine* e, const long* steps, int nsteps) {
    return guard([&] {
        for (int i = 0; i < nsteps; i++) {
            const long n = (long)e->E->slots();
            if (((steps[i] % n) + n) % n) e->E->galois_key(e->E->galois_for_rotation(steps[i]));
        }
    });
}
int ckks_set_keys_external(ckks_engine* e, int external) {
    e->E->keys_external = external != 0;
    return CKKS_OK;
}
int ckks_switch_key_ids(ckks_engine* e, uint64_t* ids_out, int capacity, int* count) {
    return guard([&] {
        std::vector<u64> ids = e->E->switch_key_ids();
        *count = (int)ids.size();
        for (int i = 0; i < (int)ids.size() && i < capacity; i++) ids_out[i] = ids[i];
    });
}
int ckks_switch_key_buffer(ckks_engine* e, uint64_t id, void** ptr, size_t* bytes) {
    return guard([&] {
        size_t words = 0;
        u64* p = e->E->switch_key_buffer(id, &words);
        if (!p) throw std::runtime_error("no such switching key");
        e->E->sync();
        *ptr = p;
        *bytes = words * sizeof(u64);
    });
}
int ckks_set_bootstrap_params(ckks_engine* e, int K, int deg, int r, int cts, int stc) {
    return guard([&] {
        if (e->E->boot) throw std::runtime_error("bootstrap key already created");
        BootParams& b = e->E->prm.boot;
        b.K = K; b.cheb_degree = deg; b.double_angle = r; b.cts_groups = cts; b.stc_groups = stc;
    });
}
int ckks_keygen_bootstrap(ckks_engine* e) { return guard([&] { e->E->bootstrap_setup(); }); }

int ckks_encode(ckks_engine* e, const double* z, int level, ckks_pt** out) {
    return guard([&] { *out = reinterpret_cast<ckks_pt*>(e->E->encode(z, level)); });
}
int ckks_encrypt(ckks_engine* e, const double* z, int level, ckks_ct** out) {
    return guard([&] { *out = H(e->E->encrypt(z, level)); });
}
int ckks_decrypt(ckks_engine* e, const ckks_ct* ct, double* z) { return guard([&] { e->E->decrypt(C(ct), z); }); }
int ckks_encrypt_zeta16(ckks_engine* e, const uint8_t* nibbles, int level, ckks_ct** out) {
    return guard([&] { *out = H(e->E->encrypt_zeta16(nibbles, level)); });
}
int ckks_decrypt_zeta16(ckks_engine* e, const ckks_ct* ct, uint8_t* nibbles_out) {
    return guard([&] { e->E->decrypt_zeta16(C(ct), nibbles_out); });
}
int ckks_snap_zeta16(ckks_engine* e, const ckks_ct* ct, int level, int stride, ckks_ct** out) {
    return guard([&] { *out = H(e->E->snap_zeta16(C(ct), level, stride)); });
}
int ckks_ct_batch(const ckks_ct* ct) { return C(ct)->nb; }
int ckks_encrypt_batch(ckks_engine* e, const double* z, int nb, int level, ckks_ct** out) {
    return guard([&] { *out = H(e->E->encrypt(z, level, nb)); });
}
int ckks_encrypt_zeta16_batch(ckks_engine* e, const uint8_t* nibbles, int nb, int level, ckks_ct** out) {
    return guard([&] { *out = H(e->E->encrypt_zeta16(nibbles, level, nb)); });
}
int ckks_ct_stack(ckks_engine* e, ckks_ct* const* items, int n, ckks_ct** out) {
    return guard([&] {
        std::vector<Ct*> v(n);
        for (int i = 0; i < n; i++) v[i] = C(items[i]);
        *out = H(e->E->stack(v));
    });
}
int ckks_ct_slice(ckks_engine* e, const ckks_ct* ct, int start, int count, ckks_ct** out) {
    return guard([&] { *out = H(e->E->slice(C(ct), start, count)); });
}
int ckks_ct_item(ckks_engine* e, const ckks_ct* ct, int index, ckks_ct** out) {
    return guard([&] { *out = H(e->E->item(C(ct), index)); });
}
void ckks_ct_free(ckks_engine* e, ckks_ct* ct) { e->E->free_ct(C(ct)); }
void ckks_pt_free(ckks_engine* e, ckks_pt* pt) { e->E->free_pt(reinterpret_cast<Pt*>(pt)); }
int ckks_ct_level(const ckks_ct* ct) { return C(ct)->level; }
int ckks_ct_npoly(const ckks_ct* ct) { return C(ct)->npoly; }
int ckks_pt_level(const ckks_pt* pt) { return P(pt)->level; }

int ckks_add(ckks_engine* e, ckks_ct* a, ckks_ct* b, ckks_ct** out) { return guard([&] { *out = H(e->E->add(C(a), C(b))); }); }
int ckks_sub(ckks_engine* e, ckks_ct* a, ckks_ct* b, ckks_ct** out) { return guard([&] { *out = H(e->E->sub(C(a), C(b))); }); }
int ckks_negate(ckks_engine* e, const ckks_ct* a, ckks_ct** out) { return guard([&] { *out = H(e->E->negate(C(a))); }); }
int ckks_mul(ckks_engine* e, ckks_ct* a, ckks_ct* b, ckks_ct** out) { return guard([&] { *out = H(e->E->mul(C(a), C(b))); }); }
int ckks_mul_norelin(ckks_engine* e, ckks_ct* a, ckks_ct* b, ckks_ct** out) {
    return guard([&] { *out = H(e->E->mul_norelin(C(a), C(b))); });
}
int ckks_relinearize(ckks_engine* e, const ckks_ct* a, ckks_ct** out) { return guard([&] { *out = H(e->E->relinearize(C(a))); }); }
int ckks_mul_const(ckks_engine* e, const ckks_ct* a, double re, double im, ckks_ct** out) {
    return guard([&] { *out = H(e->E->mul_const(C(a), re, im)); });
}
int ckks_mul_plain(ckks_engine* e, const ckks_ct* a, const ckks_pt* p, ckks_ct** out) {
    return guard([&] { *out = H(e->E->mul_plain(C(a), P(p))); });
}
int ckks_add_const(ckks_engine* e, const ckks_ct* a, double re, double im, ckks_ct** out) {
    return guard([&] { *out = H(e->E->add_const(C(a), re, im)); });
}
int ckks_add_plain(ckks_engine* e, const ckks_ct* a, const ckks_pt* p, ckks_ct** out) {
    return guard([&] { *out = H(e->E->add_plain(C(a), P(p))); });
}
int ckks_mul_i(ckks_engine* e, const ckks_ct* a, int sign, ckks_ct** out) { return guard([&] { *out = H(e->E->mul_i(C(a), sign)); }); }
int ckks_level_down(ckks_engine* e, ckks_ct* a, int level, ckks_ct** out) {
    return guard([&] { *out = H(e->E->copy(e->E->level_down(C(a), level))); });
}
int ckks_power_basis(ckks_engine* e, ckks_ct* a, int degree, ckks_ct** out) {
    return guard([&] {
        std::vector<Ct*> v = e->E->power_basis(C(a), degree);
        for (int i = 0; i < degree; i++) out[i] = H(v[i]);
    });
}
int ckks_power_basis_sparse(ckks_engine* e, ckks_ct* a, int degree, const int* exponents, int n, ckks_ct** out) {
    return guard([&] {
        std::vector<unsigned char> need(degree > 0 ? degree : 1, 0);
        for (int i = 0; i < n; i++) {
            if (exponents[i] < 1 || exponents[i] > degree) throw std::runtime_error("make_power_basis: exponent out of range");
            need[exponents[i] - 1] = 1;
        }
        std::vector<Ct*> v = e->E->power_basis(C(a), degree, need.data());
        for (int i = 0; i < degree; i++) out[i] = v[i] ? H(v[i]) : nullptr;
    });
}
int ckks_conjugate(ckks_engine* e, const ckks_ct* a, ckks_ct** out) { return guard([&] { *out = H(e->E->conjugate(C(a))); }); }
int ckks_rotate(ckks_engine* e, const ckks_ct* a, long steps, ckks_ct** out) {
    return guard([&] { *out = H(e->E->rotate(C(a), steps)); });
}
int ckks_rotate_hoisted(ckks_engine* e, const ckks_ct* a, const long* steps, int nsteps, ckks_ct** out) {
    return guard([&] {
        std::vector<Ct*> v = e->E->rotate_hoisted(C(a), std::vector<long>(steps, steps + nsteps));
        for (int i = 0; i < nsteps; i++) out[i] = H(v[i]);
    });
}
int ckks_lut2(ckks_engine* e, ckks_ct* const* A, ckks_ct* const* B, int nbasis, const int* p, const int* q,
              const double* coef, int nterms, ckks_ct** out) {
    return guard([&] {
        std::vector<Ct*> a(nbasis), b(nbasis);
        for (int i = 0; i < nbasis; i++) { a[i] = C(A[i]); b[i] = C(B[i]); }
        *out = H(e->E->lut2(a, b, p, q, coef, nterms));
    });
}
int ckks_lincomb(ckks_engine* e, ckks_ct* const* X, int n, const double* coef, ckks_ct** out) {
    return guard([&] {
        std::vector<Ct*> x(n);
        for (int i = 0; i < n; i++) x[i] = C(X[i]);
        *out = H(e->E->lincomb(x, coef, n));
    });
}
int ckks_bootstrap(ckks_engine* e, ckks_ct* a, ckks_ct** out) { return guard([&] { *out = H(e->E->bootstrap(C(a))); }); }
int ckks_bootstrap_out_level(const ckks_engine* e) { return e->E->boot_out_level(); }

int ckks_counters(const ckks_engine* e, long* o) {
    o[0] = e->E->n_keyswitch; o[1] = e->E->n_ntt_limbs; o[2] = e->E->n_rescale; o[3] = e->E->n_mul_cc; o[4] = e->E->n_boot;
    return CKKS_OK;
}

int ckks_arena_stats(const ckks_engine* e, long* driver_allocs, size_t* arena_bytes, size_t* cached_bytes) {
    *driver_allocs = e->E->n_driver_allocs;
    *arena_bytes = e->E->driver_bytes;
    *cached_bytes = 0;
    for (auto& kv : e->E->arenas)
        for (int i = 0; i < Engine::kMaxLanes; i++) *cached_bytes += kv.second->pools[i].cached;
    return CKKS_OK;
}

// ---------------------------------------------------------------- raw access (tests)
int ckks_ct_export(ckks_engine* e, const ckks_ct* ct, uint64_t* out) {
    return guard([&] {
        const Ct* c = C(ct);
        dev::d2h(out, c->d, (size_t)c->nb * c->npoly * (c->level + 1) * e->E->N() * sizeof(u64), e->E->st);
        e->E->sync();
    });
}
int ckks_ct_import(ckks_engine* e, int npoly, int level, const uint64_t* data, ckks_ct** out) {
    return guard([&] {
        Ct* c = e->E->new_ct(npoly, level);
        dev::h2d(c->d, data, (size_t)npoly * (level + 1) * e->E->N() * sizeof(u64), e->E->st);
        e->E->sync();
        *out = H(c);
    });
}
int ckks_ct_import_batch(ckks_engine* e, int nb, int npoly, int level, const uint64_t* data, ckks_ct** out) {
    return guard([&] {
        Ct* c = e->E->new_ct(npoly, level, nb);
        dev::h2d(c->d, data, (size_t)nb * npoly * (level + 1) * e->E->N() * sizeof(u64), e->E->st);
        e->E->sync();
        *out = H(c);
    });
}
int ckks_pt_export(ckks_engine* e, const ckks_pt* pt, uint64_t* out) {
    return guard([&] {
        dev::d2h(out, P(pt)->d, (size_t)(P(pt)->level + 1) * e->E->N() * sizeof(u64), e->E->st);
        e->E->sync();
    });
}
int ckks_export_secret(ckks_engine* e, int64_t* out) {
    return guard([&] {
        const std::vector<i64>& s = e->E->sk_host();
        if (s.empty()) throw std::runtime_error("no secret key");
        for (size_t i = 0; i < s.size(); i++) out[i] = s[i];
    });
}
int ckks_export_public(ckks_engine* e, uint64_t* out) {
    return guard([&] {
        if (!e->E->has_pk) throw std::runtime_error("no public key");
        dev::d2h(out, e->E->pk_dev(), (size_t)2 * (e->E->L() + 1) * e->E->N() * sizeof(u64), e->E->st);
        e->E->sync();
    });
}
int ckks_export_switch_key(ckks_engine* e, uint64_t g, uint64_t* out) {
    return guard([&] {
        const EvalKey* k = g ? e->E->galois_key(g) : e->E->relin_key();
        if (!k->d) throw std::runtime_error("no such key");
        dev::d2h(out, k->d, (size_t)e->E->dnum() * 2 * e->E->nmod() * e->E->N() * sizeof(u64), e->E->st);
        e->E->sync();
    });
}
int ckks_test_ntt(ckks_engine* e, uint64_t* data, int nrows, const int* mods, int inverse) {
    return guard([&] {
        Engine& E = *e->E;
        const size_t n = E.N();
        std::vector<int> done;
        // rows are processed in chunks of at most CKKS_MAX_MODULI
        for (int r0 = 0; r0 < nrows; r0 += CKKS_MAX_MODULI) {
            const int nr = std::min(CKKS_MAX_MODULI, nrows - r0);
            u64* d = E.alloc((size_t)nr * n);
            dev::h2d(d, data + (size_t)r0 * n, (size_t)nr * n * sizeof(u64), E.st);
            std::vector<int> rows(nr), m(mods + r0, mods + r0 + nr);
            for (int i = 0; i < nr; i++) rows[i] = i;
            E.ntt_rows(d, rows, m, inverse != 0);
            dev::d2h(data + (size_t)r0 * n, d, (size_t)nr * n * sizeof(u64), E.st);
            E.sync();
            E.release(d);
        }
    });
}
int ckks_test_automorph(ckks_engine* e, uint64_t* data, int nrows, uint64_t g) {
    return guard([&] {
        Engine& E = *e->E;
        const size_t n = E.N();
        u64* a = E.alloc((size_t)nrows * n);
        u64* b = E.alloc((size_t)nrows * n);
        dev::h2d(a, data, (size_t)nrows * n * sizeof(u64), E.st);
        E.automorph(b, a, nrows, 1, g);
        dev::d2h(data, b, (size_t)nrows * n * sizeof(u64), E.st);
        E.sync();
        E.release(a);
        E.release(b);
    });
}
int ckks_test_key_switch(ckks_engine* e, const uint64_t* poly, int level, uint64_t g, uint64_t* out) {
    return guard([&] {
        Engine& E = *e->E;
        const size_t n = E.N(), ps = (size_t)(level + 1) * n;
        const EvalKey* k = g ? E.galois_key(g) : E.relin_key();
        if (!k->d) throw std::runtime_error("no such key");
        u64* a = E.alloc(ps);
        u64* o = E.alloc(2 * ps);
        dev::h2d(a, poly, ps * sizeof(u64), E.st);
        E.key_switch(a, level, k, o);
        dev::d2h(out, o, 2 * ps * sizeof(u64), E.st);
        E.sync();
        E.release(a);
        E.release(o);
    });
}
uint64_t ckks_galois_for_rotation(const ckks_engine* e, long steps) { return e->E->galois_for_rotation(steps); }

// ---------------------------------------------------------------- timing
// ---------------------------------------------------------------- captured graphs
int ckks_graph_create(ckks_engine* e, int* id) { return guard([&] { *id = e->E->graph_create(); }); }
int ckks_graph_enter(ckks_engine* e, int id) { return guard([&] { e->E->graph_enter(id); }); }
int ckks_graph_leave(ckks_engine* e) { return guard([&] { e->E->graph_leave(); }); }
int ckks_graph_capture_begin(ckks_engine* e, int id) { return guard([&] { e->E->graph_capture_begin(id); }); }
int ckks_graph_capture_end(ckks_engine* e, int id) { return guard([&] { e->E->graph_capture_end(id); }); }
int ckks_graph_capture_abort(ckks_engine* e) { return guard([&] { e->E->graph_capture_abort(); }); }
int ckks_graph_launch(ckks_engine* e, int id, int replay_stream) { return guard([&] { e->E->graph_launch(id, replay_stream); }); }
int ckks_graph_wait(ckks_engine* e, int replay_stream) { return guard([&] { e->E->graph_wait(replay_stream); }); }
int ckks_graph_destroy(ckks_engine* e, int id) { return guard([&] { e->E->graph_destroy(id); }); }
int ckks_graph_info(ckks_engine* e, int id, long* nodes, long* launches, size_t* arena_bytes, long* capture_misses) {
    return guard([&] {
        Engine::GraphRec* G = e->E->graph_rec(id);
        if (nodes) *nodes = G->g ? (long)G->g->nodes : 0;
        if (launches) *launches = G->launches;
        if (arena_bytes) *arena_bytes = G->arena.bytes;
        if (capture_misses) *capture_misses = G->arena.capture_misses;
    });
}
int ckks_ct_assign(ckks_engine* e, ckks_ct* dst, const ckks_ct* src, int replay_stream) {
    return guard([&] { e->E->ct_assign(C(dst), C(src), replay_stream); });
}
int ckks_ct_clear_memo(ckks_engine* e, ckks_ct* ct) { return guard([&] { e->E->ct_clear_memo(C(ct)); }); }

int ckks_timer_start(ckks_engine* e) { return guard([&] { e->timer.start(e->E->st); }); }
int ckks_timer_stop_ms(ckks_engine* e, float* ms) { return guard([&] { *ms = e->timer.stop_ms(e->E->st); }); }

int ckks_profile_ntt_begin(ckks_engine* e) { return guard([&] { e->E->profile_begin(); }); }
int ckks_profile_ntt_end(ckks_engine* e, double* ms, long* calls, long* limbs) {
    return guard([&] { e->E->profile_end(ms, calls, limbs); });
}
int ckks_bench_ntt(ckks_engine* e, int nlimbs, int batches, int inverse, int iters, float* ms_out) {
    return guard([&] {
        Engine& E = *e->E;
        const size_t n = E.N();
        if (nlimbs > E.nmod()) throw std::runtime_error("bench_ntt: too many limbs");
        u64* d = E.alloc((size_t)batches * nlimbs * n);
        std::vector<int> rows(nlimbs);
        for (int i = 0; i < nlimbs; i++) rows[i] = i;
        for (int b = 0; b < batches; b++)
            launch_sample_uniform(E.ks, d + (size_t)b * nlimbs * n, E.limb_list(rows), 1234, 77 + b, E.st);
        for (int w = 0; w < 3; w++) E.ntt_rows(d, rows, rows, inverse != 0, batches, (size_t)nlimbs * n);
        e->timer.start(E.st);
        for (int it = 0; it < iters; it++) E.ntt_rows(d, rows, rows, inverse != 0, batches, (size_t)nlimbs * n);
        *ms_out = e->timer.stop_ms(E.st) / iters;
        E.release(d);
    });
}
static Ct* random_ct(Engine& E, int level, u64 tag, int nb = 1) {
    Ct* c = E.new_ct(2, level, nb);
    LimbList ll = E.limb_list(E.mods_q(level));
    const size_t ps = (size_t)(level + 1) * E.N();
    for (int k = 0; k < 2 * nb; k++) launch_sample_uniform(E.ks, c->d + k * ps, ll, 99, tag + k, E.st);
    return c;
}
int ckks_bench_rotate_batch(ckks_engine* e, int level, int nb, int iters, float* ms_out) {
    return guard([&] {
        Engine& E = *e->E;
        Ct* c = random_ct(E, level, 10, nb);
        const long step = (long)E.slots() / 4;
        for (int w = 0; w < 3; w++) E.free_ct(E.rotate(c, step));
        e->timer.start(E.st);
        for (int it = 0; it < iters; it++) E.free_ct(E.rotate(c, step));
        *ms_out = e->timer.stop_ms(E.st) / iters;
        E.free_ct(c);
    });
}
int ckks_bench_mul_batch(ckks_engine* e, int level, int nb, int iters, float* ms_out) {
    return guard([&] {
        Engine& E = *e->E;
        Ct* a = random_ct(E, level, 20, nb);
        Ct* b = random_ct(E, level, 300, nb);
        for (int w = 0; w < 3; w++) E.free_ct(E.mul(a, b));
        e->timer.start(E.st);
        for (int it = 0; it < iters; it++) E.free_ct(E.mul(a, b));
        *ms_out = e->timer.stop_ms(E.st) / iters;
        E.free_ct(a);
        E.free_ct(b);
    });
}
int ckks_bench_rotate(ckks_engine* e, int level, int iters, float* ms_out) {
    return guard([&] {
        Engine& E = *e->E;
        Ct* c = random_ct(E, level, 10);
        const long step = (long)E.slots() / 4;
        for (int w = 0; w < 3; w++) E.free_ct(E.rotate(c, step));
        e->timer.start(E.st);
        for (int it = 0; it < iters; it++) E.free_ct(E.rotate(c, step));
        *ms_out = e->timer.stop_ms(E.st) / iters;
        E.free_ct(c);
    });
}
// the same with `lanes` independent ciphertexts rotated concurrently on stream lanes (throughput rather than latency)
int ckks_bench_rotate_lanes(ckks_engine* e, int level, int lanes, int iters, float* ms_per_rotation) {
    return guard([&] {
        Engine& E = *e->E;
        if (lanes < 1 || lanes > 16) throw std::runtime_error("bench_rotate_lanes: 1..16 lanes");
        std::vector<Ct*> cts;
        for (int l = 0; l < lanes; l++) cts.push_back(random_ct(E, level, 40 + 2 * l));
        const long step = (long)E.slots() / 4;
        auto round = [&](int reps) {
            if (lanes > 1) E.fork(lanes);
            for (int l = 0; l < lanes; l++) {
                if (lanes > 1) E.set_lane(l);
                for (int it = 0; it < reps; it++) E.free_ct(E.rotate(cts[l], step));
            }
            if (lanes > 1) E.join();
        };
        round(3);
        e->timer.start(E.st);
        round(iters);
        *ms_per_rotation = e->timer.stop_ms(E.st) / (iters * lanes);
        for (Ct* c : cts) E.free_ct(c);
    });
}
int ckks_bench_mul(ckks_engine* e, int level, int iters, float* ms_out) {
    return guard([&] {
        Engine& E = *e->E;
        Ct* a = random_ct(E, level, 20);
        Ct* b = random_ct(E, level, 30);
        for (int w = 0; w < 3; w++) E.free_ct(E.mul(a, b));
        e->timer.start(E.st);
        for (int it = 0; it < iters; it++) E.free_ct(E.mul(a, b));
        *ms_out = e->timer.stop_ms(E.st) / iters;
        E.free_ct(a);
        E.free_ct(b);
    });
}

}  // extern "C"
