// engine.cu -- host orchestration of the CKKS engine: parameters, tables, keys, encode/encrypt/
// decrypt, arithmetic with canonical-scale level alignment, hybrid key switching, Galois maps.
// Restated independently (same written spec) by oracle/ckks_oracle.py, against
// which every integer result here is checked bit-for-bit.
#include "engine.cuh"

#include <math.h>

#include <algorithm>
#include <set>

long g_launch_count = 0;
#ifdef CKKS_TIME_LAUNCHES
double g_launch_host_ns = 0;
#endif
#ifdef CKKS_EMU
thread_local emu_uint3 blockIdx, threadIdx;
thread_local dim3 blockDim, gridDim;
namespace emu { std::vector<std::function<void()>>* g_record = nullptr; }
#endif
namespace dev { int g_capturing = 0; }

namespace ckks {

typedef unsigned __int128 u128;

// ------------------------------------------------------------------ host modular helpers
u64 mulmod_h(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }
u64 powmod_h(u64 a, u64 e, u64 q) {
    u64 r = 1 % q;
    a %= q;
    while (e) {
        if (e & 1) r = mulmod_h(r, a, q);
        a = mulmod_h(a, a, q);
        e >>= 1;
    }
    return r;
}
u64 invmod_h(u64 a, u64 q) { return powmod_h(a % q, q - 2, q); }
u64 shoup_h(u64 w, u64 q) { return (u64)((((u128)w) << 64) / q); }
static u64 mix64_h(u64 z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
u64 rand64_h(u64 seed, u64 stream, u64 idx) { return mix64_h(mix64_h(seed + stream * 0xD1342543DE82EF95ull) + idx); }
static u64 bitrev_h(u64 x, int bits) {
    u64 r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
static u64 stream_id(u64 kind, u64 a = 0, u64 b = 0) { return (kind << 48) | (a << 16) | b; }
enum { ST_SK = 1, ST_PK_A, ST_PK_E, ST_EVK_A, ST_EVK_E, ST_ENC_V, ST_ENC_E0, ST_ENC_E1 };

// ------------------------------------------------------------------ prime chain (spec S1)
static bool is_prime_h(u64 n) {
    if (n < 2) return false;
    static const u64 small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (u64 p : small)
        if (n % p == 0) return n == p;
    u64 d = n - 1;
    int r = 0;
    while ((d & 1) == 0) { d >>= 1; r++; }
    for (u64 a : small) {   // deterministic for n < 3.3e24
        u64 x = powmod_h(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int i = 0; i < r - 1; i++) {
            x = mulmod_h(x, x, n);
            if (x == n - 1) { comp = false; break; }
        }
        if (comp) return false;
    }
    return true;
}
static std::vector<u64> primes_below(u64 bound, u64 step, int count, std::set<u64>& used) {
    std::vector<u64> out;
    u64 c = (bound - 2) / step * step + 1;
    while ((int)out.size() < count) {
        if (!used.count(c) && is_prime_h(c)) { out.push_back(c); used.insert(c); }
        c -= step;
    }
    return out;
}
static u64 prime_nearest(double target, u64 step, std::set<u64>& used) {
    const i64 k0 = (i64)nearbyint((target - 1.0) / (double)step);
    for (i64 d = 0; d < (1 << 20); d++) {
        for (int s = 0; s < (d ? 2 : 1); s++) {
            const i64 k = s == 0 ? k0 - d : k0 + d;
            if (k <= 0) continue;
            const u64 c = (u64)k * step + 1;
            if (c > 2 && !used.count(c) && is_prime_h(c)) { used.insert(c); return c; }
        }
    }
    throw std::runtime_error("no prime found");
}
// Message ratio q_0 / S_0 of the chain (ModRaise, and the head-room the modulus-256 XOR outputs of the reference's table
// need at level 0: SURVEY.md H3).  A chain whose q_0 is SMALLER than 2^(scale_bits + ratio) keeps the ratio by letting the
// scale descend: S_l = 2^(scale_bits - drop 2^-l), drop = scale_bits + ratio - q0_bits.  The deviation from 2^scale_bits
// doubles per level downwards (S_{l-1} = S_l^2 / q_l), so every q_l still sits next to 2^scale_bits, the levels >= 6 carry
// 2^scale_bits within a fifth of a bit, and only S_2, S_1, S_0 give up 2.5, 5 and 10 bits (the XOR4 outputs that land
// there tolerate 4e-3).  With q0_bits = 50 the whole chain, q_0 included, is below the FP64 limit: no kernel takes the
// 64-bit integer path any more.  drop = 0 (q0_bits = scale_bits + ratio, the 60-bit q_0) is the uniform chain.
static const int kRatioBits = 10;
static double chain_drop(int scale_bits, int q0_bits) { return std::max(0, scale_bits + kRatioBits - q0_bits); }
static double desired_scale(int scale_bits, double drop, int l) {
    return drop > 0 ? pow(2.0, (double)scale_bits - drop * ldexp(1.0, -l)) : ldexp(1.0, scale_bits);
}
static std::vector<double> canonical_scales(const std::vector<u64>& q, double top_scale) {
    const int L = (int)q.size() - 1;
    std::vector<double> s(L + 1);
    s[L] = top_scale;
    for (int l = L; l >= 1; l--) s[l - 1] = s[l] * s[l] / (double)q[l];
    return s;
}

// top_levels > 0: the highest `top_levels` levels carry the larger scale 2^top_bits (CoeffToSlot runs there: its error is
// governed by the plaintext / rescale precision relative to the huge ModRaise values); the scale then descends to
// 2^scale_bits as fast as primes below 2^60.5 allow.  top_levels = 0 is the uniform chain.
Params default_params(int logn, int levels, int scale_bits, int q0_bits, int p_bits, int dnum, int hamming,
                      int fresh_level, int top_levels, int top_bits) {
    Params P;
    P.logn = logn;
    P.scale_bits = scale_bits;
    P.top_levels = top_levels;
    P.top_bits = top_bits;
    P.hamming = hamming;
    const u64 step = 2ull << logn;
    std::set<u64> used;
    const int nq = levels + 1;
    P.alpha = (nq + dnum - 1) / dnum;
    // the special primes (largest below 2^p_bits) never collide with the chain (which stays below 2^60.5 < 2^p_bits - ..):
    // reserve a generous set first so `used` protects them, then cut it down to the K the widest digit needs
    std::vector<u64> reserve = primes_below(1ull << p_bits, step, 16, used);
    P.q.assign(nq, 0);
    P.q[0] = primes_below(1ull << q0_bits, step, 1, used)[0];
    const double hi = ldexp(1.0, top_bits), cap = pow(2.0, 60.5);
    const double drop = chain_drop(scale_bits, q0_bits);
    P.top_scale = top_levels > 0 ? hi : desired_scale(scale_bits, drop, levels);
    P.scale_drop = (int)drop;
    double s = P.top_scale;
    for (int l = levels; l >= 1; l--) {
        const double delta = desired_scale(scale_bits, drop, l - 1);
        const double want = (top_levels > 0 && l - 1 > levels - top_levels) ? hi : delta;   // desired S_{l-1}
        const double next = std::max(want, s * s / cap);
        const double target = s * s / next;
        P.q[l] = prime_nearest(target, step, used);
        s = s * s / (double)P.q[l];
    }
    // P must dominate the widest key-switch digit (alpha consecutive limbs): sum of the limbs' bit lengths
    int digit_bits = 0;
    for (int j = 0; j * P.alpha < nq; j++) {
        int bits = 0;
        for (int i = j * P.alpha; i < std::min((j + 1) * P.alpha, nq); i++) bits += 64 - __builtin_clzll(P.q[i]);
        digit_bits = std::max(digit_bits, bits);
    }
    const int K = (digit_bits + 1 + (p_bits - 1) - 1) / (p_bits - 1);
    P.p.assign(reserve.begin(), reserve.begin() + K);
    P.fresh_level = fresh_level < 0 ? levels : std::min(fresh_level, levels);
    return P;
}

// ------------------------------------------------------------------ construction
static u64 find_psi(u64 q, int logn) {
    const u64 twoN = 2ull << logn, N = 1ull << logn;
    for (u64 x = 2;; x++) {
        const u64 r = powmod_h(x, (q - 1) / twoN, q);
        if (powmod_h(r, N, q) == q - 1) return r;
    }
}

template <typename T>
static T* upload(Engine* E, const std::vector<T>& h, std::vector<void*>& owned) {
    T* d = (T*)dev::alloc(h.size() * sizeof(T), E->st);
    dev::h2d(d, h.data(), h.size() * sizeof(T), E->st);
    dev::sync(E->st);
    owned.push_back(d);
    return d;
}

Engine::Engine(const Params& P) : prm(P) {
    if (prm.logn != 16 && prm.logn != 12) throw std::runtime_error("engine: logn must be 16 (or 12 for tests)");
    if (prm.q.empty() || prm.p.empty()) throw std::runtime_error("engine: empty modulus chain");
    dev::set_device(prm.device);
    dev::pool_setup(prm.device);
    st = streams[0] = dev::stream_create();
    lane_made[0] = lane_busy[0] = true;
    arenas[0] = &main_arena;
    if (const char* v = getenv("CKKS_NTT_FUSE")) fuse_ntt = atoi(v) != 0;
    if (const char* v = getenv("CKKS_TENSOR_FUSE")) fuse_tensor = atoi(v) != 0;
    if (const char* v = getenv("CKKS_KS_ADD_FUSE")) fuse_ks_add = atoi(v) != 0;
    if (const char* v = getenv("CKKS_MUL_FACTOR_FUSE")) fuse_mul_factor = atoi(v) != 0;
    if (const char* v = getenv("CKKS_ALIGN_FUSE")) fuse_align = atoi(v) != 0;
    if (const char* v = getenv("CKKS_BC_MMA")) bc_mode = atoi(v) != 0 ? BC_MMA : bc_mode;
    if (const char* v = getenv("CKKS_BC_FP")) bc_mode = atoi(v) != 0 ? BC_FP : (bc_mode == BC_FP ? BC_INT : bc_mode);
    if (const char* v = getenv("CKKS_BC_FP_INT_EVERY")) bc_fp_int_every = atoi(v);
    int ntt_cluster = 0;                                                                 // DESIGN.md 8.1; a field of THIS engine's tables
    if (const char* v = getenv("CKKS_NTT_CLUSTER")) ntt_cluster = atoi(v);
    if (const char* v = getenv("CKKS_CHEB_DEGREE")) prm.boot.cheb_degree = atoi(v);      // tuning / A-B runs only
    mod = prm.q;
    mod.insert(mod.end(), prm.p.begin(), prm.p.end());
    if ((int)mod.size() > CKKS_MAX_MODULI) throw std::runtime_error("engine: too many moduli");
    if (prm.alpha > BC_MAX_SRC || K() > BC_MAX_SRC) throw std::runtime_error("engine: digit too wide");
    if (L() + 1 + K() > BC_MAX_TGT) throw std::runtime_error("engine: too many conversion targets");
    if ((dnum()) > NTT_MAX_Z) throw std::runtime_error("engine: dnum too large");
    scales = canonical_scales(prm.q, prm.top_scale > 0 ? prm.top_scale
                                                        : ldexp(1.0, prm.top_levels > 0 ? prm.top_bits : prm.scale_bits));
    const size_t n = N();
    const int nm = nmod();
    std::vector<ModConst> mc(nm);
    std::vector<u64> fwd((size_t)nm * n), fwd_s((size_t)nm * n), inv((size_t)nm * n), inv_s((size_t)nm * n);
    std::vector<double> fwd_d((size_t)nm * n), inv_d((size_t)nm * n);
    psi.resize(nm);
    Jroot.resize(nm);
    for (int i = 0; i < nm; i++) {
        const u64 q = mod[i];
        if (q >> 61) throw std::runtime_error("engine: moduli must be below 2^61");
        if ((q - 1) % (2 * n)) throw std::runtime_error("engine: modulus is not 1 mod 2N");
        psi[i] = find_psi(q, prm.logn);
        Jroot[i] = powmod_h(psi[i], n / 2, q);
        const u64 ipsi = invmod_h(psi[i], q);
        u64 pw = 1, ipw = 1;
        u64 *F = &fwd[(size_t)i * n], *Fs = &fwd_s[(size_t)i * n], *I = &inv[(size_t)i * n], *Is = &inv_s[(size_t)i * n];
        for (size_t e = 0; e < n; e++) {
            const size_t k = bitrev_h(e, prm.logn);
            F[k] = pw; Fs[k] = shoup_h(pw, q);
            I[k] = ipw; Is[k] = shoup_h(ipw, q);
            fwd_d[(size_t)i * n + k] = (double)pw;
            inv_d[(size_t)i * n + k] = (double)ipw;
            pw = mulmod_h(pw, psi[i], q);
            ipw = mulmod_h(ipw, ipsi, q);
        }
        ModConst& m = mc[i];
        m.q = q;
        int k = 64 - __builtin_clzll(q);
        m.k1 = k - 1;
        m.pad = 0;
        m.mu = (u64)((((u128)1) << (k + 63)) / q);
        m.ninv = invmod_h(n % q, q);
        m.ninv_s = shoup_h(m.ninv, q);
        m.w1n = mulmod_h(I[1], m.ninv, q);
        m.w1n_s = shoup_h(m.w1n, q);
    }
    d_fwd = upload(this, fwd, owned);
    d_fwd_s = upload(this, fwd_s, owned);
    d_inv = upload(this, inv, owned);
    d_inv_s = upload(this, inv_s, owned);
    d_mc = upload(this, mc, owned);
    tabs = NttTables{d_fwd, d_fwd_s, d_inv, d_inv_s, upload(this, fwd_d, owned), upload(this, inv_d, owned), d_mc,
                     prm.logn, ntt_cluster};
    ks = KShape{d_mc, prm.logn};
    // canonical-embedding tables (spec S9)
    const size_t M = 2 * n, ns = n / 2;
    std::vector<u32> rot(ns);
    u64 pw = 1;
    for (size_t j = 0; j < ns; j++) { rot[j] = (u32)pw; pw = (pw * 5) % M; }
    std::vector<double> ksi(2 * (M + 1));
    for (size_t k = 0; k <= M; k++) {
        const double ang = 2.0 * M_PI * (double)k / (double)M;
        ksi[2 * k] = cos(ang);
        ksi[2 * k + 1] = sin(ang);
    }
    // per-level scalar tables: P^-1 mod q_i, q_l^-1 mod q_i (i < l)
    {
        std::vector<int> qi = mods_q(L());
        std::vector<u64> pinv(L() + 1);
        for (int i = 0; i <= L(); i++) {
            u64 pp = 1;
            for (u64 pk_ : prm.p) pp = mulmod_h(pp, pk_ % mod[i], mod[i]);
            pinv[i] = invmod_h(pp, mod[i]);
        }
        scalar_list(pinv, qi, sl_pinv);
        std::vector<u64> pm(L() + 1);
        for (int i = 0; i <= L(); i++) {
            u64 pp = 1;
            for (u64 pk_ : prm.p) pp = mulmod_h(pp, pk_ % mod[i], mod[i]);
            pm[i] = pp;
        }
        scalar_list(pm, qi, sl_pmodq);
        sl_qinv.resize(L() + 1);
        for (int l = 1; l <= L(); l++) {
            std::vector<int> lo = mods_q(l - 1);
            std::vector<u64> inv(l);
            for (int i = 0; i < l; i++) inv[i] = invmod_h(mod[l] % mod[i], mod[i]);
            scalar_list(inv, lo, sl_qinv[l]);
        }
    }
    {
        std::vector<double> z16(32);
        for (int k = 0; k < 16; k++) {
            z16[2 * k] = cos(-2.0 * M_PI * k / 16.0);
            z16[2 * k + 1] = sin(-2.0 * M_PI * k / 16.0);
        }
        d_zeta16 = upload(this, z16, owned);
    }
    d_rot = upload(this, rot, owned);
    d_ksi = upload(this, ksi, owned);
    std::vector<int> flag(1, 0);
    d_flag = upload(this, flag, owned);
    std::vector<u64> epoch(1, 0);
    d_epoch = upload(this, epoch, owned);
}

Engine::~Engine() {
    try {
        if (dev::capturing()) graph_capture_abort();
        while (in_fork()) join();
        sync();
        arena = &main_arena;
        while (!graphs.empty()) graph_destroy(graphs.begin()->first);
    } catch (...) {}
    bootstrap_teardown();
    for (auto& kv : gkeys) dev::free(kv.second.d, st);
    for (auto& kv : perms) dev::free(kv.second, st);
    for (auto& kv : const_tabs) dev::free(kv.second, st);
    for (auto& kv : modup_dev) dev::free(kv.second, st);
    for (auto& kv : moddown_dev) dev::free(kv.second, st);
    for (auto& kv : index_tabs) dev::free(kv.second, st);
    dev::free(relin.d, st);
    dev::free(sk_ntt, st);
    dev::free(pk, st);
    for (void* p : owned) dev::free(p, st);
    try { trim_pools(); } catch (...) {}
    try { sync(); } catch (...) {}
    for (int i = 0; i < kMaxLanes; i++)
        if (lane_made[i]) dev::stream_destroy(streams[i]);
    for (int i = 1; i < kMaxReplay; i++)
        if (replay[i]) dev::stream_destroy(replay[i]);
}

// ------------------------------------------------------------------ lanes
void Engine::sync() {
    for (int i = 0; i < kMaxLanes; i++)
        if (lane_made[i]) dev::sync(streams[i]);
}
void Engine::fork(int k) {
    if (k < 1) throw std::runtime_error("fork: need at least one lane");
    Frame F;
    F.parent = cur_lane;
    F.epoch = ++epoch_counter;
    if (!lanes_on) {                       // serial mode: every "lane" is the parent stream itself
        F.lanes.assign(k, cur_lane);
        F.serial = true;
        frames.push_back(std::move(F));
        return;
    }
    for (int i = 1; i < kMaxLanes && (int)F.lanes.size() < k; i++) {
        if (lane_busy[i]) continue;
        if (!lane_made[i]) { streams[i] = dev::stream_create(); lane_made[i] = true; }
        lane_busy[i] = true;
        F.lanes.push_back(i);
    }
    if ((int)F.lanes.size() < k) {
        for (int l : F.lanes) lane_busy[l] = false;
        throw std::runtime_error("fork: out of stream lanes");
    }
    // every lane is ordered after the parent stream from here on, and the parent enqueues nothing until the join, so a
    // lane may take cached buffers from the pools of its ancestors (see alloc); sibling pools stay private
    for (int l : F.lanes) dev::stream_wait(streams[l], streams[F.parent]);
    frames.push_back(std::move(F));
    set_lane(0);
}
void Engine::set_lane(int i) {
    if (frames.empty()) throw std::runtime_error("set_lane outside fork/join");
    Frame& F = frames.back();
    if (i < 0 || i >= (int)F.lanes.size()) throw std::runtime_error("set_lane: no such lane in this fork");
    cur_lane = F.lanes[i];
    st = streams[cur_lane];
}
void Engine::join() {
    if (frames.empty()) return;
    Frame F = std::move(frames.back());
    frames.pop_back();
    // the parent is ordered after every lane: their cached buffers return to the parent's pool
    LanePool* pools = arena->pools;
    LanePool& PP = pools[F.parent];
    if (F.serial) F.lanes.clear();
    for (int l : F.lanes) {
        dev::stream_wait(streams[F.parent], streams[l]);
        for (auto& kv : pools[l].free) {
            std::vector<void*>& dst = PP.free[kv.first];
            dst.insert(dst.end(), kv.second.begin(), kv.second.end());
            PP.cached += kv.first * kv.second.size();
            kv.second.clear();
        }
        pools[l].cached = 0;
        lane_busy[l] = false;
    }
    cur_lane = F.parent;
    st = streams[cur_lane];
    for (Ct* c : F.dct) free_ct(c);        // re-examined in the enclosing frame (may be deferred again)
    for (Pt* p : F.dpt) free_pt(p);
}

// ------------------------------------------------------------------ memory
// size classes: whole limbs (N words), rounded up along 1, 2, 3, 4, 6, 8, 12, 16, 24, 32, ... so that the many
// slightly different buffer shapes of a key switch recycle each other's memory (at most 1/3 slack)
static size_t size_class_limbs(size_t limbs) {
    size_t c = 1;
    while (c < limbs) {
        const size_t mid = c + c / 2;
        if (c >= 2 && mid >= limbs) return mid;
        c *= 2;
    }
    return c;
}
u64* Engine::alloc(size_t words) {
    const size_t limbs = (words + N() - 1) / N();
    const size_t bytes = size_class_limbs(limbs ? limbs : 1) * N() * sizeof(u64);
    // own pool first, then the pools of the enclosing frames' parent streams (idle and ordered before this lane)
    LanePool* pools = arena->pools;
    int lane = cur_lane;
    for (int depth = (int)frames.size(); depth >= 0; depth--) {
        LanePool& P = pools[lane];
        auto it = P.free.find(bytes);
        if (it != P.free.end() && !it->second.empty()) {
            void* p = it->second.back();
            it->second.pop_back();
            P.cached -= bytes;
            return (u64*)p;
        }
        if (depth == 0) break;
        lane = frames[depth - 1].parent;
    }
    // miss: stream-ordered driver memory, or -- while a graph is being captured -- plain device memory owned by the
    // graph's arena (dev::alloc switches by itself)
    void* p = dev::alloc(bytes, st);
    alloc_bytes[p] = Buf{bytes, arena->id};
    if (dev::capturing()) { arena->plain.push_back(p); arena->capture_misses++; }
    n_driver_allocs++;
    driver_bytes += bytes;
    arena->bytes += bytes;
    return (u64*)p;
}
// a buffer always returns to the arena it was born in: the scratch a captured graph replays into never becomes the
// memory of an eager operation, and the other way round
void Engine::release(void* p) {
    if (!p) return;
    auto it = alloc_bytes.find(p);
    if (it == alloc_bytes.end()) { dev::free(p, st); return; }
    const Buf b = it->second;
    if (b.arena != arena->id) {
        auto ia = arenas.find(b.arena);
        if (ia == arenas.end()) {                 // its arena is gone (graph destroyed): back to the driver
            alloc_bytes.erase(it);
            driver_bytes -= b.bytes;
            if (!dev::capturing()) dev::free(p, st);
            return;
        }
        LanePool& P = ia->second->pools[0];
        P.free[b.bytes].push_back(p);
        P.cached += b.bytes;
        return;
    }
    LanePool& P = arena->pools[cur_lane];
    P.free[b.bytes].push_back(p);
    P.cached += b.bytes;
    if (P.cached > ((size_t)48 << 30) && !dev::capturing()) trim_pools();
}
// give everything cached in the current arena back to the driver pool (only when a lane hoards more than 48 GiB, and at
// teardown); plain allocations made during a capture stay with their arena until it is destroyed
void Engine::trim_pools() {
    sync();
    std::set<void*> plain(arena->plain.begin(), arena->plain.end());
    for (LanePool& P : arena->pools) {
        for (auto& kv : P.free) {
            std::vector<void*> keep;
            for (void* p : kv.second) {
                if (plain.count(p)) { keep.push_back(p); continue; }
                alloc_bytes.erase(p);
                dev::free(p, streams[0]);
                driver_bytes -= kv.first;
                arena->bytes -= kv.first;
                P.cached -= kv.first;
            }
            kv.second.swap(keep);
        }
    }
}

// ------------------------------------------------------------------ captured graphs
// A Graph owns a private arena.  Life cycle (driven through the C ABI by desilofhe.Engine.capture):
//   graph_create -> graph_enter (the arena becomes current) -> [static inputs are copied, one eager warm-up run fills the
//   arena and creates every lazily built table] -> graph_capture_begin -> the same calls again, now recorded ->
//   graph_capture_end (instantiate) -> graph_leave.  graph_launch replays it on a replay stream; graph_wait orders the
//   engine's main stream after that replay.
int Engine::graph_create() {
    if (in_fork()) throw std::runtime_error("graph_create inside a fork");
    GraphRec* G = new GraphRec();
    G->id = ++graph_counter;
    G->arena.id = G->id;
    graphs[G->id] = G;
    arenas[G->id] = &G->arena;
    return G->id;
}
Engine::GraphRec* Engine::graph_rec(int id) {
    auto it = graphs.find(id);
    if (it == graphs.end()) throw std::runtime_error("no such graph");
    return it->second;
}
void Engine::graph_enter(int id) {
    if (in_fork()) throw std::runtime_error("graph_enter inside a fork");
    if (arena != &main_arena) throw std::runtime_error("another graph arena is already current");
    arena = &graph_rec(id)->arena;
}
void Engine::graph_leave() {
    if (in_fork()) throw std::runtime_error("graph_leave inside a fork");
    if (dev::capturing()) throw std::runtime_error("graph_leave during a capture");
    arena = &main_arena;
}
void Engine::graph_capture_begin(int id) {
    GraphRec* G = graph_rec(id);
    if (arena != &G->arena) throw std::runtime_error("graph_capture_begin: enter the graph's arena first");
    if (in_fork()) throw std::runtime_error("graph_capture_begin inside a fork");
    if (G->g) throw std::runtime_error("graph already captured");
    sync();                                          // everything the capture reads has been produced
    G->launches0 = g_launch_count;
    G->cnt0[0] = n_keyswitch; G->cnt0[1] = n_ntt_limbs; G->cnt0[2] = n_rescale; G->cnt0[3] = n_mul_cc; G->cnt0[4] = n_boot;
    capture_id = G->id;
    dev::capture_begin(streams[0]);
    // first node of every graph: advance the replay epoch, so the encryptions recorded in the graph (device-side hard
    // renorm) draw fresh randomness on every replay instead of the values frozen at capture time
    launch_bump(d_epoch, streams[0]);
}
void Engine::graph_capture_end(int id) {
    GraphRec* G = graph_rec(id);
    if (!dev::capturing() || capture_id != id) throw std::runtime_error("graph_capture_end without a matching begin");
    if (in_fork()) { dev::capture_abort(streams[0]); capture_id = 0; throw std::runtime_error("graph_capture_end inside a fork"); }
    capture_id = 0;
    G->g = dev::capture_end(streams[0]);
    G->launches = g_launch_count - G->launches0;
    const long now[5] = {n_keyswitch, n_ntt_limbs, n_rescale, n_mul_cc, n_boot};
    for (int i = 0; i < 5; i++) G->cnt[i] = now[i] - G->cnt0[i];
}
void Engine::graph_capture_abort() {
    while (in_fork()) { try { join(); } catch (...) { break; } }
    dev::capture_abort(streams[0]);
    capture_id = 0;
}
dev_stream Engine::replay_stream(int slot) {
    if (slot < 0 || slot >= kMaxReplay) throw std::runtime_error("replay stream index out of range");
    if (slot == 0) return streams[0];
    if (!replay[slot]) replay[slot] = dev::stream_create();
    return replay[slot];
}
void Engine::graph_launch(int id, int slot) {
    GraphRec* G = graph_rec(id);
    if (!G->g) throw std::runtime_error("graph has not been captured");
    if (in_fork() || dev::capturing()) throw std::runtime_error("graph_launch inside a fork or capture");
    dev_stream s = replay_stream(slot);
    if (slot) dev::stream_wait(s, streams[0]);      // its inputs were produced on the main stream
    dev::graph_launch(G->g, s);
    g_launch_count += G->launches;
    n_keyswitch += G->cnt[0]; n_ntt_limbs += G->cnt[1]; n_rescale += G->cnt[2]; n_mul_cc += G->cnt[3]; n_boot += G->cnt[4];
    G->replays++;
}
void Engine::graph_wait(int slot) {
    if (slot) dev::stream_wait(streams[0], replay_stream(slot));
}
// overwrite the contents of a static input of a graph (same shape) on the replay stream the graph will be launched on
void Engine::ct_assign(Ct* dst, const Ct* src, int slot) {
    if (dst->npoly != src->npoly || dst->level != src->level || dst->nb != src->nb)
        throw std::runtime_error("ct_assign: shape mismatch");
    if (dst->d == src->d) return;                  // the static input itself: nothing to copy
    dev_stream s = replay_stream(slot);
    if (slot) dev::stream_wait(s, streams[0]);
    dev::d2d(dst->d, src->d, (size_t)src->nb * src->npoly * (src->level + 1) * N() * sizeof(u64), s);
    if (slot) dev::stream_wait(streams[0], s);      // src may be released (and rewritten) by the main stream afterwards
}
void Engine::ct_clear_memo(Ct* c) {
    for (auto& kv : c->lowered) free_ct(kv.second);
    c->lowered.clear();
}
void Engine::graph_destroy(int id) {
    GraphRec* G = graph_rec(id);
    if (arena == &G->arena) throw std::runtime_error("graph_destroy: leave the graph's arena first");
    sync();
    for (int i = 1; i < kMaxReplay; i++)
        if (replay[i]) dev::sync(replay[i]);
    dev::graph_destroy(G->g);
    std::set<void*> plain(G->arena.plain.begin(), G->arena.plain.end());
    for (LanePool& P : G->arena.pools)
        for (auto& kv : P.free)
            for (void* p : kv.second) {
                alloc_bytes.erase(p);
                driver_bytes -= kv.first;
                if (plain.count(p)) { dev::free_plain(p); plain.erase(p); }
                else dev::free(p, streams[0]);
            }
    // plain buffers still held by live ciphertexts are left to the process teardown (they are few: the graph's outputs)
    arenas.erase(id);
    graphs.erase(id);
    delete G;
}
Ct* Engine::new_ct(int npoly, int level, int nb) {
    if (nb < 1) throw std::runtime_error("new_ct: batch size must be positive");
    Ct* c = new Ct();
    c->npoly = npoly;
    c->level = level;
    c->nb = nb;
    c->lane = cur_lane;
    c->epoch = cur_epoch();
    c->cap = capture_id;
    c->d = alloc((size_t)nb * npoly * (level + 1) * N());
    return c;
}
int Engine::batch_of(const Ct* a, const Ct* b) const {
    if (a->nb != b->nb && a->nb != 1 && b->nb != 1)
        throw std::runtime_error("operands hold different batch sizes (" + std::to_string(a->nb) + " and " +
                                 std::to_string(b->nb) + ")");
    return std::max(a->nb, b->nb);
}
Ct* Engine::stack(const std::vector<Ct*>& items) {
    if (items.empty()) throw std::runtime_error("stack: no ciphertexts");
    const Ct* f = items[0];
    int nb = 0;
    for (const Ct* c : items) {
        if (!c || c->npoly != f->npoly || c->level != f->level)
            throw std::runtime_error("stack: items must be ciphertexts of one shape");
        nb += c->nb;
    }
    Ct* r = new_ct(f->npoly, f->level, nb);
    const size_t per = (size_t)f->npoly * (f->level + 1) * N();
    size_t at = 0;
    for (const Ct* c : items) {                       // a batched item contributes all its items, in order
        dev::d2d(r->d + at * per, c->d, (size_t)c->nb * per * sizeof(u64), st);
        at += c->nb;
    }
    return r;
}
Ct* Engine::slice(const Ct* c, int start, int count) {
    if (start < 0 || count < 1 || start + count > c->nb) throw std::runtime_error("slice: batch range out of bounds");
    Ct* r = new_ct(c->npoly, c->level, count);
    const size_t per = (size_t)c->npoly * (c->level + 1) * N();
    dev::d2d(r->d, c->d + (size_t)start * per, (size_t)count * per * sizeof(u64), st);
    return r;
}
Ct* Engine::item(const Ct* c, int i) {
    if (i < 0 || i >= c->nb) throw std::runtime_error("item: batch index out of range");
    return slice(c, i, 1);
}
void Engine::free_ct(Ct* c) {
    if (!c) return;
    if (in_fork() && !(c->lane == cur_lane && c->epoch == cur_epoch())) {
        frames.back().dct.push_back(c);            // another lane may still be reading it: released at the join
        return;
    }
    for (auto& kv : c->lowered) free_ct(kv.second);
    release(c->d);
    delete c;
}
void Engine::free_pt(Pt* p) {
    if (!p) return;
    if (in_fork()) { frames.back().dpt.push_back(p); return; }
    release(p->d);
    delete p;
}

// ------------------------------------------------------------------ limb helpers
std::vector<int> Engine::mods_q(int level) const {
    std::vector<int> v(level + 1);
    for (int i = 0; i <= level; i++) v[i] = i;
    return v;
}
std::vector<int> Engine::mods_qp(int level) const {
    std::vector<int> v = mods_q(level);
    for (int k = 0; k < K(); k++) v.push_back(L() + 1 + k);
    return v;
}
LimbList Engine::limb_list(const std::vector<int>& mods) const {
    LimbList l;
    memset(&l, 0, sizeof(l));
    l.n = (int)mods.size();
    for (int i = 0; i < l.n; i++) l.idx[i] = (unsigned char)mods[i];
    return l;
}
void Engine::scalar_list(const std::vector<u64>& vals, const std::vector<int>& mods, ScalarList& out) const {
    memset(&out, 0, sizeof(out));
    for (size_t i = 0; i < mods.size(); i++) {
        const u64 q = mod[mods[i]];
        out.v[i] = vals[i] % q;
        out.vs[i] = shoup_h(out.v[i], q);
    }
}
static u64 signed_residue(i64 x, u64 q) {
    const u64 r = (u64)(x < 0 ? -(u128)(i64)x : (u128)x) % q;   // |x| < 2^63
    return (x < 0 && r) ? q - r : r;
}
void Engine::const_residues(double re, double im, double scale, const std::vector<int>& mods, ScalarList& cp,
                            ScalarList& cm) const {
    const double a = nearbyint(re * scale), b = nearbyint(im * scale);
    if (!(fabs(a) < 9.0e18) || !(fabs(b) < 9.0e18)) throw std::runtime_error("constant does not fit 63 bits at this scale");
    const i64 R = (i64)a, I = (i64)b;
    std::vector<u64> vp(mods.size()), vm(mods.size());
    for (size_t i = 0; i < mods.size(); i++) {
        const u64 q = mod[mods[i]];
        const u64 r = signed_residue(R, q), ij = mulmod_h(signed_residue(I, q), Jroot[mods[i]], q);
        vp[i] = (r + ij) % q;
        vm[i] = (r + q - ij) % q;
    }
    scalar_list(vp, mods, cp);
    scalar_list(vm, mods, cm);
}

void Engine::ntt_rows(u64* data, const std::vector<int>& rows, const std::vector<int>& mods, bool inverse, int nz,
                      size_t zstride, int nb, size_t bstride) {
    if (rows.empty() || nz == 0) return;
    if (nz > NTT_MAX_Z) {   // split long batches
        for (int z0 = 0; z0 < nz; z0 += NTT_MAX_Z)
            ntt_rows(data + z0 * zstride, rows, mods, inverse, std::min(NTT_MAX_Z, nz - z0), zstride, nb, bstride);
        return;
    }
    NttJob J;
    memset(&J, 0, sizeof(J));
    J.n = (int)rows.size();
    J.nz = nz;
    J.nb = nb;
    J.sbs = J.dbs = bstride;
    J.szs = J.dzs = zstride;
    for (int z = 0; z < nz; z++)
        for (int i = 0; i < J.n; i++) {
            J.rows[z][i] = J.srows[z][i] = (unsigned char)rows[i];
            J.mods[z][i] = (unsigned char)mods[i];
        }
    run_ntt(data, data, J, inverse, (long)J.n * nz * nb);
}

// every NTT launch goes through here: limb accounting and, when profiling, a CUDA-event pair per call on the
// engine's stream (bench.py's roofline leg: average duration and algorithmic bytes of the NTT kernels)
void Engine::run_ntt(const u64* src, u64* dst, const NttJob& J, bool inverse, long limbs, const u64* src2) {
    dev::Timer* t = nullptr;
    if (prof_on) {
        if (prof_used == prof_timers.size()) prof_timers.emplace_back();
        t = &prof_timers[prof_used++];
        t->start(st);
    }
    if (inverse) ntt_inverse(src, dst, J, tabs, st, src2);
    else ntt_forward(src, dst, J, tabs, st);
    if (t) { t->mark_stop(st); prof_limbs += limbs; prof_calls++; }
    n_ntt_limbs += limbs;
}
// forward transform with the rescale-lift prologue and / or the (a - NTT(x)) * s epilogue fused in (ntt.cuh: NttFuse)
void Engine::run_ntt_fused(const u64* src, u64* dst, const NttJob& J, const NttFuse& F, long limbs) {
    dev::Timer* t = nullptr;
    if (prof_on) {
        if (prof_used == prof_timers.size()) prof_timers.emplace_back();
        t = &prof_timers[prof_used++];
        t->start(st);
    }
    ntt_forward_fused(src, dst, J, tabs, F, st);
    if (t) { t->mark_stop(st); prof_limbs += limbs; prof_calls++; }
    n_ntt_limbs += limbs;
}
void Engine::profile_begin() {
    prof_on = true;
    prof_used = 0;
    prof_limbs = prof_calls = 0;
    prof_ms = 0;
}
void Engine::profile_end(double* ms, long* calls, long* limbs) {
    dev::sync(st);
    double total = prof_ms;
    for (size_t i = 0; i < prof_used; i++) total += prof_timers[i].elapsed_ms();
    prof_on = false;
    *ms = total; *calls = prof_calls; *limbs = prof_limbs;
}

// ------------------------------------------------------------------ Galois maps (spec S4)
u64 Engine::galois_for_rotation(long steps) const {
    const long n = (long)slots();
    long r = steps % n;
    if (r < 0) r += n;
    const u64 e = (u64)((n - r) % n);
    return powmod_h(5, e, 2 * N());
}
const u32* Engine::galois_perm(u64 g) {
    auto it = perms.find(g);
    if (it != perms.end()) return it->second;
    const size_t n = N(), M = 2 * n;
    std::vector<u32> h(n);
    for (size_t k = 0; k < n; k++) {
        const u64 e = (2 * bitrev_h(k, prm.logn) + 1) * g % M;
        h[k] = (u32)bitrev_h((e - 1) / 2, prm.logn);
    }
    u32* d = (u32*)dev::alloc(n * sizeof(u32), st);
    dev::h2d(d, h.data(), n * sizeof(u32), st);
    dev::sync(st);
    perms[g] = d;
    return d;
}
void Engine::automorph(u64* out, const u64* in, int rows, int npoly, u64 g) {
    const size_t stride = (size_t)rows * N();
    launch_permute(ks, out, in, galois_perm(g), rows, npoly, PolyStride{stride, stride, 0}, st);
}
void Engine::automorph(u64* out, const u64* in, int rows, int npoly, u64 g, PolyStride ps) {
    launch_permute(ks, out, in, galois_perm(g), rows, npoly, ps, st);
}

// ------------------------------------------------------------------ keys (spec S8)
void Engine::keygen_secret() {
    const size_t n = N();
    const int h = prm.hamming;
    std::vector<u32> perm(n);
    sk_coef.assign(n, 0);
    for (size_t k = 0; k < n; k++) perm[k] = (u32)k;
    const u64 sid = stream_id(ST_SK);
    for (int i = 0; i < h; i++) {
        const u64 j = i + rand64_h(prm.seed, sid, (u64)i) % (n - i);
        std::swap(perm[i], perm[j]);
        sk_coef[perm[i]] = (rand64_h(prm.seed, sid, (u64)h + i) & 1) ? 1 : -1;
    }
    i64* d_s = (i64*)dev::alloc(n * sizeof(i64), st);
    dev::h2d(d_s, sk_coef.data(), n * sizeof(i64), st);
    std::vector<int> all(nmod());
    for (int i = 0; i < nmod(); i++) all[i] = i;
    if (!sk_ntt) sk_ntt = alloc((size_t)nmod() * n);
    launch_reduce_i64(ks, sk_ntt, d_s, limb_list(all), st);
    ntt_rows(sk_ntt, all, all, false);
    dev::sync(st);
    dev::free(d_s, st);
    has_sk = true;
}

void Engine::keygen_public() {
    if (!has_sk) throw std::runtime_error("public key needs a secret key");
    const size_t n = N();
    const int nq = L() + 1;
    std::vector<int> idx = mods_q(L());
    LimbList ll = limb_list(idx);
    if (!pk) pk = alloc((size_t)2 * nq * n);
    u64* b = pk;
    u64* a = pk + (size_t)nq * n;
    u64* e = alloc((size_t)nq * n);
    launch_sample_uniform(ks, a, ll, prm.seed, stream_id(ST_PK_A), st);
    launch_sample_small(ks, e, ll, prm.seed, stream_id(ST_PK_E), 0, st);
    ntt_rows(e, idx, idx, false);
    PolyStride z{0, 0, 0};
    launch_mul(ks, b, a, sk_ntt, ll, 1, z, st);         // a*s  (sk rows 0..L line up with q rows)
    launch_sub(ks, b, e, b, ll, 1, z, st);              // e - a*s
    release(e);
    has_pk = true;
}

// evk[j] = (-a_j s + e_j + P * F_j * s_from,  a_j) over Q_L u P; F_j = CRT selector of digit j
EvalKey Engine::make_switch_key(u64 key_id, const u64* s_from_ntt) {
    if (!has_sk) throw std::runtime_error("switching key needs a secret key");
    const size_t n = N();
    const int rows = nmod();
    std::vector<int> idx = mods_qp(L());           // == 0..nmod-1
    LimbList ll = limb_list(idx);
    const int dn = dnum();
    EvalKey key;
    key.d = alloc((size_t)dn * 2 * rows * n);
    if (keys_external) return key;                 // content arrives by broadcast from the rank that generated it
    u64* e = alloc((size_t)rows * n);
    u64* t = alloc((size_t)rows * n);
    PolyStride z{0, 0, 0};
    for (int j = 0; j < dn; j++) {
        u64* b = key.d + ((size_t)j * 2) * rows * n;
        u64* a = b + (size_t)rows * n;
        launch_sample_uniform(ks, a, ll, prm.seed, stream_id(ST_EVK_A, key_id, j), st);
        launch_sample_small(ks, e, ll, prm.seed, stream_id(ST_EVK_E, key_id, j), 0, st);
        ntt_rows(e, idx, idx, false);
        launch_mul(ks, b, a, sk_ntt, ll, 1, z, st);
        launch_sub(ks, b, e, b, ll, 1, z, st);
        std::vector<u64> fac(rows, 0);
        for (int i = j * prm.alpha; i < std::min((j + 1) * prm.alpha, L() + 1); i++) {
            u64 pp = 1;
            for (u64 pk_ : prm.p) pp = mulmod_h(pp, pk_ % mod[i], mod[i]);
            fac[i] = pp;
        }
        ScalarList sc;
        scalar_list(fac, idx, sc);
        launch_mul_scalar(ks, t, s_from_ntt, ll, sc, 1, z, st);
        launch_add(ks, b, b, t, ll, 1, z, st);
    }
    release(e);
    release(t);
    return key;
}

void Engine::keygen_relin() {
    const size_t n = N();
    std::vector<int> idx = mods_qp(L());
    u64* s2 = alloc((size_t)nmod() * n);
    launch_mul(ks, s2, sk_ntt, sk_ntt, limb_list(idx), 1, PolyStride{0, 0, 0}, st);
    if (relin.d) release(relin.d);
    relin = make_switch_key(0, s2);
    release(s2);
    has_relin = true;
}

std::vector<u64> Engine::switch_key_ids() const {
    std::vector<u64> ids;
    if (relin.d) ids.push_back(0);
    for (auto& kv : gkeys) ids.push_back(kv.first);
    return ids;
}
u64* Engine::switch_key_buffer(u64 id, size_t* words) {
    *words = (size_t)dnum() * 2 * nmod() * N();
    if (id == 0) return relin.d;
    auto it = gkeys.find(id);
    return it == gkeys.end() ? nullptr : it->second.d;
}

EvalKey* Engine::galois_key(u64 g) {
    auto it = gkeys.find(g);
    if (it != gkeys.end()) return &it->second;
    u64* sg = alloc((size_t)nmod() * N());
    automorph(sg, sk_ntt, nmod(), 1, g);
    EvalKey k = make_switch_key(g, sg);
    release(sg);
    dev::sync(st);                                  // first use may come from the other lane
    gkeys[g] = k;
    return &gkeys[g];
}

// ------------------------------------------------------------------ encode / encrypt / decrypt (spec S9, S10)
// polynomial strides + batch strides of a launch (kernels.cuh: PolyStride)
static PolyStride psb(size_t o, size_t a, size_t b, int nb, size_t bo, size_t ba, size_t bb) {
    PolyStride p{o, a, b};
    p.nb = nb; p.bout = bo; p.ba = ba; p.bb = bb;
    return p;
}
// z_dev: nb x 2n doubles on the device (overwritten as scratch) -> nb x N signed coefficients
void Engine::encode_coeffs_from_dev(i64* out_dev, double* z_dev, double scale, bool check, int nb) {
    const size_t ns = slots();
    double* w = (double*)alloc((size_t)nb * 2 * ns);
    launch_special_ifft(ks, w, z_dev, d_rot, d_ksi, st, nb);
    launch_round_coeffs(ks, out_dev, w, scale, d_flag, st, nb);
    release(w);
    if (!check) return;
    int flag = 0;
    dev::d2h(&flag, d_flag, sizeof(int), st);
    dev::sync(st);
    if (flag) {
        int zero = 0;
        dev::h2d(d_flag, &zero, sizeof(int), st);
        dev::sync(st);
        throw std::runtime_error("plaintext coefficient does not fit 62 bits");
    }
}
void Engine::encode_coeffs_dev(i64* out_dev, const double* z_host, double scale, int nb) {
    const size_t ns = slots();
    double* z = (double*)alloc((size_t)nb * 2 * ns);
    dev::h2d(z, z_host, (size_t)nb * 2 * ns * sizeof(double), st);
    try { encode_coeffs_from_dev(out_dev, z, scale, true, nb); } catch (...) { release(z); throw; }
    release(z);
}

Pt* Engine::encode(const double* z, int level) {
    if (level < 0 || level > L()) throw std::runtime_error("encode: bad level");
    const size_t n = N();
    i64* coef = (i64*)alloc(n);
    try { encode_coeffs_dev(coef, z, scales[level]); } catch (...) { release(coef); throw; }
    Pt* p = new Pt();
    p->level = level;
    p->d = alloc((size_t)(level + 1) * n);
    std::vector<int> idx = mods_q(level);
    launch_reduce_i64(ks, p->d, coef, limb_list(idx), st);
    ntt_rows(p->d, idx, idx, false);
    release(coef);
    return p;
}

Ct* Engine::encrypt(const double* z, int level, int nb) {
    if (!has_pk) throw std::runtime_error("encrypt needs a public key");
    if (level < 0) level = prm.fresh_level;
    if (level > L()) throw std::runtime_error("encrypt: bad level");
    if (nb < 1) throw std::runtime_error("encrypt: batch size must be positive");
    i64* coef = (i64*)alloc((size_t)nb * N());
    try { encode_coeffs_dev(coef, z, scales[level], nb); } catch (...) { release(coef); throw; }
    Ct* c = encrypt_coeffs(coef, level, nb);
    release(coef);
    return c;
}

// public-key encryption of nb x N signed message coefficients already on the device (spec S8, S10).  Item i of a batch
// draws the streams an unbatched encryption number (counter + i) would draw: a batched encryption is bit-identical to
// nb consecutive single ones.
Ct* Engine::encrypt_coeffs(const i64* coef, int level, int nb) {
    const size_t n = N();
    const int nl = level + 1, nq = L() + 1;
    std::vector<int> idx = mods_q(level);
    LimbList ll = limb_list(idx);
    const u64 k = enc_counter;
    enc_counter += (u64)nb;
    // per item: t[0] = v, t[1] = e0 + m, t[2] = e1  (coefficient domain), one batched NTT for all of them
    const size_t ps = (size_t)nl * n, tb = 3 * ps;
    u64* t = alloc((size_t)nb * tb);
    launch_sample_small(ks, t, ll, prm.seed, stream_id(ST_ENC_V, k), 1, st, nb, tb, d_epoch);
    launch_sample_small(ks, t + ps, ll, prm.seed, stream_id(ST_ENC_E0, k), 0, st, nb, tb, d_epoch);
    launch_sample_small(ks, t + 2 * ps, ll, prm.seed, stream_id(ST_ENC_E1, k), 0, st, nb, tb, d_epoch);
    u64* m = alloc((size_t)nb * ps);
    launch_reduce_i64(ks, m, coef, ll, st, nb, ps);
    launch_add(ks, t + ps, t + ps, m, ll, 1, psb(0, 0, 0, nb, tb, tb, ps), st);
    ntt_rows(t, idx, idx, false, 3, ps, nb, tb);
    Ct* c = new_ct(2, level, nb);
    // c_k = v * pk_k + t[k+1]
    launch_mul(ks, c->d, t, pk, ll, 2, psb(ps, 0, (size_t)nq * n, nb, 2 * ps, tb, 0), st);
    launch_add(ks, c->d, c->d, t + ps, ll, 2, psb(ps, ps, ps, nb, 2 * ps, 2 * ps, tb), st);
    release(t);
    release(m);
    return c;
}

// secret-key decryption to slot values left ON THE DEVICE (nb x 2n doubles, caller releases)
double* Engine::decrypt_to_dev(const Ct* c) {
    if (!has_sk) throw std::runtime_error("decrypt needs a secret key");
    // spec S10: limb 0 alone carries the message when |m| S_l < q_0 / 2.  On a descending-scale chain (q_0 next to the scale
    // primes) that only holds at level 0: align to level 0 first
    if (prm.scale_drop > 0 && c->level > 0) {
        Ct* low = lowered_copy(c, 0);              // not memoised on c: a graph's output handle is rewritten by the next replay
        double* zz = decrypt_to_dev(low);
        free_ct(low);
        return zz;
    }
    const size_t n = N(), ns = slots();
    const int nb = c->nb;
    const size_t ps = (size_t)(c->level + 1) * n, cb = (size_t)c->npoly * ps;    // polynomial / batch strides of c
    std::vector<int> idx{0};
    LimbList ll = limb_list(idx);
    PolyStride z0{0, 0, 0};
    // t = c0 + c1 s (+ c2 s^2) on limb 0 (spec S10)
    u64* t = alloc((size_t)nb * n);
    u64* tmp = alloc((size_t)nb * n);
    u64* sp = nullptr;
    for (int k = 1; k < c->npoly; k++) {
        const u64* spow = sk_ntt;
        if (k > 1) {
            if (!sp) { sp = alloc(n); launch_mul(ks, sp, sk_ntt, sk_ntt, ll, 1, z0, st); }
            else launch_mul(ks, sp, sp, sk_ntt, ll, 1, z0, st);
            spow = sp;
        }
        launch_mul(ks, tmp, c->d + k * ps, spow, ll, 1, psb(0, 0, 0, nb, n, cb, 0), st);
        launch_add(ks, t, k == 1 ? c->d : t, tmp, ll, 1, psb(0, 0, 0, nb, n, k == 1 ? cb : n, n), st);
    }
    ntt_rows(t, idx, idx, true, 1, 0, nb, n);
    double* w = (double*)alloc((size_t)nb * 2 * ns);
    double* zz = (double*)alloc((size_t)nb * 2 * ns);
    launch_center_to_w(ks, w, t, 0, scales[c->level], st, nb);
    launch_special_fft(ks, zz, w, d_rot, d_ksi, st, nb);
    release(t); release(tmp);
    if (sp) release(sp);
    release(w);
    return zz;
}

// hard renorm without leaving the device (reference pipeline.py:65-69 does decrypt -> host snap -> encrypt):
// decrypt, snap every slot to the nearest zeta_16 codeword, re-encrypt at `level`
Ct* Engine::snap_zeta16(const Ct* a, int level, int stride) {
    if (!has_pk) throw std::runtime_error("renorm needs a public key");
    if (level < 0) level = prm.fresh_level;
    if (level > L()) throw std::runtime_error("renorm: bad level");
    double* zz = decrypt_to_dev(a);
    launch_snap_zeta16(ks, zz, d_zeta16, stride < 1 ? 1 : stride, st, a->nb);
    i64* coef = (i64*)alloc((size_t)a->nb * N());
    encode_coeffs_from_dev(coef, zz, scales[level], false, a->nb);       // unit-modulus slots cannot overflow
    Ct* c = encrypt_coeffs(coef, level, a->nb);
    release(coef);
    release(zz);
    return c;
}

// zeta_16 codec on the device: the caller ships one nibble per slot (n bytes instead of 16 n), the codeword lookup,
// the embedding and the encryption run here; decryption returns the nearest-codeword index per slot (the host side of
// reference state_encoder.py:14-38 / utils.py:9-19 moved next to the data)
Ct* Engine::encrypt_zeta16(const unsigned char* nib_host, int level, int nb) {
    if (!has_pk) throw std::runtime_error("encrypt needs a public key");
    if (level < 0) level = prm.fresh_level;
    if (level > L()) throw std::runtime_error("encrypt: bad level");
    if (nb < 1) throw std::runtime_error("encrypt: batch size must be positive");
    const size_t ns = slots();
    unsigned char* nib = (unsigned char*)alloc(((size_t)nb * ns + 7) / 8);
    double* z = (double*)alloc((size_t)nb * 2 * ns);
    i64* coef = (i64*)alloc((size_t)nb * N());
    dev::h2d(nib, nib_host, (size_t)nb * ns, st);
    launch_zeta16_from_nibbles(ks, z, nib, d_zeta16, st, nb);
    encode_coeffs_from_dev(coef, z, scales[level], false, nb);        // unit-modulus slots cannot overflow
    Ct* c = encrypt_coeffs(coef, level, nb);
    dev::sync(st);                                                    // nib_host may be reused by the caller
    release(coef); release(z); release(nib);
    return c;
}
void Engine::decrypt_zeta16(const Ct* c, unsigned char* nib_out_host) {
    const size_t ns = slots();
    double* zz = decrypt_to_dev(c);
    unsigned char* nib = (unsigned char*)alloc(((size_t)c->nb * ns + 7) / 8);
    launch_nibbles_from_zeta16(ks, nib, zz, st, c->nb);
    dev::d2h(nib_out_host, nib, (size_t)c->nb * ns, st);
    dev::sync(st);
    release(nib);
    release(zz);
}

void Engine::decrypt(const Ct* c, double* z_out) {
    double* zz = decrypt_to_dev(c);
    dev::d2h(z_out, zz, (size_t)c->nb * 2 * slots() * sizeof(double), st);
    dev::sync(st);
    release(zz);
}

// ------------------------------------------------------------------ basis conversion tables (spec S5)
BaseConvTable Engine::make_bc_table(const std::vector<int>& src, const std::vector<int>& srow,
                                    const std::vector<int>& tgt, const std::vector<int>& orow, bool exact) {
    BaseConvTable T;
    memset(&T, 0, sizeof(T));
    T.exact = exact ? 1 : 0;
    for (size_t i = 0; i < src.size(); i++) T.inv_src[i] = 1.0 / (double)mod[src[i]];
    for (size_t t = 0; t < tgt.size(); t++) {
        const u64 qt = mod[tgt[t]];
        u64 D = 1;
        for (int sidx : src) D = mulmod_h(D, mod[sidx] % qt, qt);
        T.negD[t] = (qt - D) % qt;
    }
    T.ns = (int)src.size();
    T.nt = (int)tgt.size();
    std::vector<u64> hat((size_t)T.ns * T.nt);
    for (int i = 0; i < T.ns; i++) {
        T.src[i] = (unsigned char)src[i];
        T.srow[i] = (unsigned char)srow[i];
        const u64 qi = mod[src[i]];
        u64 prod = 1;
        for (int j = 0; j < T.ns; j++)
            if (j != i) prod = mulmod_h(prod, mod[src[j]] % qi, qi);
        T.hatinv[i] = invmod_h(prod, qi);
        T.hatinv_s[i] = shoup_h(T.hatinv[i], qi);
        for (int t = 0; t < T.nt; t++) {
            const u64 qt = mod[tgt[t]];
            u64 pr = 1;
            for (int j = 0; j < T.ns; j++)
                if (j != i) pr = mulmod_h(pr, mod[src[j]] % qt, qt);
            hat[(size_t)i * T.nt + t] = pr;
        }
    }
    for (int t = 0; t < T.nt; t++) { T.tgt[t] = (unsigned char)tgt[t]; T.orow[t] = (unsigned char)orow[t]; }
    T.hat = upload(this, hat, owned);           // exact size, freed with the engine
    {
        // FP64 path (kernels.cu: k_base_convert_fp): which targets take it, which sources are split, the constants as doubles
        std::vector<double> hf((size_t)T.nt * T.ns * 2, 0.0);
        T.swide = 0;
        for (int i = 0; i < T.ns; i++)
            if (mod[src[i]] >= CKKS_FP_LIMIT) T.swide |= 1u << i;      // y_i enters as two halves below 2^32
        int eligible = 0;
        for (int t = 0; t < T.nt; t++) {
            const u64 qt = mod[tgt[t]];
            bool fp = qt < CKKS_FP_LIMIT;
            // CKKS_BC_FP_INT_EVERY=k: every k-th FP64-capable target stays on the integer pipe (pipe balance, A/B)
            if (fp && bc_fp_int_every > 0 && (++eligible % bc_fp_int_every) == 0) fp = false;
            T.tfp[t] = fp ? 1 : 0;
            T.tqinv[t] = 1.0 / (double)qt;
            T.negDd[t] = (double)T.negD[t];
            if (!fp) continue;
            for (int i = 0; i < T.ns; i++) {
                const u64 h = hat[(size_t)i * T.nt + t], h32 = mulmod_h(h, (1ull << 32) % qt, qt);
                double* e = &hf[((size_t)t * T.ns + i) * 2];
                e[0] = (double)h;
                e[1] = (double)h32;
            }
        }
        T.hatf = upload(this, hf, owned);
    }
    if (T.ns <= BC_MMA_MAX_SRC) {
        // tensor-core path (kernels.cu: k_base_convert_mma): the hat matrix cut into bytes, Toeplitz-expanded over the
        // 16 diagonals and stored in mma.m16n8k32 A-fragment order; 2^(8d) mod q_t for the recombination
        const int ntg = (T.nt + 7) / 8;
        std::vector<u32> af((size_t)ntg * 8 * 2 * 32 * 4, 0);
        auto entry = [&](int t, int d, int k) -> u32 {          // row (t, d), column k = (i, a)
            const int i = k / 8, a = k % 8, b = d - a;
            if (t >= T.nt || i >= T.ns || b < 0 || b > 7) return 0;
            return (u32)((hat[(size_t)i * T.nt + t] >> (8 * b)) & 0xff);
        };
        for (int tg = 0; tg < ntg; tg++)
            for (int j = 0; j < 8; j++)
                for (int ks = 0; ks < 2; ks++)
                    for (int lane = 0; lane < 32; lane++) {
                        const int g = lane >> 2, q4 = lane & 3, t = tg * 8 + g, k0 = ks * 32 + q4 * 4;
                        u32* r = &af[((((size_t)tg * 8 + j) * 2 + ks) * 32 + lane) * 4];
                        for (int e = 0; e < 4; e++) {
                            r[0] |= entry(t, 2 * j, k0 + e) << (8 * e);
                            r[1] |= entry(t, 2 * j + 1, k0 + e) << (8 * e);
                            r[2] |= entry(t, 2 * j, k0 + 16 + e) << (8 * e);
                            r[3] |= entry(t, 2 * j + 1, k0 + 16 + e) << (8 * e);
                        }
                    }
        std::vector<u64> p8((size_t)T.nt * 16, 0);
        for (int t = 0; t < T.nt; t++) {
            const u64 qt = mod[tgt[t]];
            u64 pw = 1 % qt;
            for (int d = 0; d < 16; d++) { p8[(size_t)t * 16 + d] = pw; pw = mulmod_h(pw, 256 % qt, qt); }
        }
        T.afrag = upload(this, af, owned);
        T.pow8 = upload(this, p8, owned);
    }
    return T;
}

// digit j of level `level`: source rows j*alpha.. in the coefficient buffer [level+1][N]; targets are
// every other modulus of Q_level u P, written to row (its position in mods_qp(level)) of ext[j]
const BaseConvTable& Engine::modup_table(int level, int digit) {
    auto key = std::make_pair(level, digit);
    auto it = modup_tabs.find(key);
    if (it != modup_tabs.end()) return it->second;
    std::vector<int> qp = mods_qp(level);
    std::vector<int> src, srow, tgt, orow;
    const int lo = digit * prm.alpha, hi = std::min((digit + 1) * prm.alpha, level + 1);
    for (int i = lo; i < hi; i++) { src.push_back(i); srow.push_back(i); }
    for (int r = 0; r < (int)qp.size(); r++)
        if (qp[r] < lo || qp[r] >= hi) { tgt.push_back(qp[r]); orow.push_back(r); }
    modup_tabs[key] = make_bc_table(src, srow, tgt, orow);
    return modup_tabs[key];
}
// {q_{level-drop+1} .. q_level} u P  ->  Q_{level-drop}: sources are rows level-drop+1..level and level+1.. of the
// accumulator, targets rows 0..level-drop of the output.  drop = 0 is the plain ModDown by P; drop = 1, 2 also
// divide by the last one or two ciphertext primes (merged rescale, spec S6b).
const BaseConvTable& Engine::moddown_table(int level, int drop) {
    const int key = level * 4 + drop;
    auto it = moddown_tabs.find(key);
    if (it != moddown_tabs.end()) return it->second;
    std::vector<int> src, srow, tgt, orow;
    for (int i = level - drop + 1; i <= level; i++) { src.push_back(i); srow.push_back(i); }
    for (int k = 0; k < K(); k++) { src.push_back(L() + 1 + k); srow.push_back(level + 1 + k); }
    for (int i = 0; i <= level - drop; i++) { tgt.push_back(i); orow.push_back(i); }
    moddown_tabs[key] = make_bc_table(src, srow, tgt, orow, true);     // exact, centred: an unbiased division by P
    return moddown_tabs[key];
}
// (P * q_{level-drop+1} .. q_level)^-1 mod q_i, i <= level - drop
const ScalarList& Engine::moddown_inv(int level, int drop, int factor) {
    const int key = (level * 4 + drop) * 4 + factor;
    auto it = moddown_invs.find(key);
    if (it != moddown_invs.end()) return it->second;
    std::vector<int> qi = mods_q(level - drop);
    std::vector<u64> inv(qi.size());
    for (size_t i = 0; i < qi.size(); i++) {
        const u64 q = mod[qi[i]];
        u64 pp = 1;
        for (u64 pk_ : prm.p) pp = mulmod_h(pp, pk_ % q, q);
        for (int j = level - drop + 1; j <= level; j++) pp = mulmod_h(pp, mod[j] % q, q);
        inv[i] = mulmod_h(invmod_h(pp, q), (u64)factor % q, q);
    }
    scalar_list(inv, qi, moddown_invs[key]);
    return moddown_invs[key];
}

const BaseConvTable* Engine::modup_tables_dev(int level) {
    auto it = modup_dev.find(level);
    if (it != modup_dev.end()) return it->second;
    const int beta = (level + 1 + prm.alpha - 1) / prm.alpha;
    std::vector<BaseConvTable> h(beta);
    for (int j = 0; j < beta; j++) h[j] = modup_table(level, j);
    BaseConvTable* d = (BaseConvTable*)dev::alloc(beta * sizeof(BaseConvTable), st);
    dev::h2d(d, h.data(), beta * sizeof(BaseConvTable), st);
    dev::sync(st);
    modup_dev[level] = d;
    return d;
}
const BaseConvTable* Engine::moddown_table_dev(int level, int drop) {
    const int key = level * 4 + drop;
    auto it = moddown_dev.find(key);
    if (it != moddown_dev.end()) return it->second;
    BaseConvTable h = moddown_table(level, drop);
    BaseConvTable* d = (BaseConvTable*)dev::alloc(sizeof(BaseConvTable), st);
    dev::h2d(d, &h, sizeof(BaseConvTable), st);
    dev::sync(st);
    moddown_dev[key] = d;
    return d;
}

// ------------------------------------------------------------------ hybrid key switching (spec S5, S6)
Decomp Engine::decompose(const u64* d, int level, const u64* times, int nb, size_t d_bs, size_t times_bs) {
    const size_t n = N();
    const int nq = level + 1, rows = nq + K();
    const int beta = (nq + prm.alpha - 1) / prm.alpha;
    const size_t eb = (size_t)beta * rows * n;             // batch stride of ext
    Decomp D;
    D.level = level;
    D.beta = beta;
    D.nb = nb;
    D.ext = alloc((size_t)nb * eb);
    D.own = d;
    D.own_bs = d_bs;
    // coefficient form of all q-limbs
    u64* coef = alloc((size_t)nb * nq * n);
    {
        NttJob J;
        memset(&J, 0, sizeof(J));
        J.n = nq; J.nz = 1;
        J.nb = nb; J.sbs = d_bs; J.dbs = (size_t)nq * n; J.s2bs = times_bs;
        for (int i = 0; i < nq; i++) { J.rows[0][i] = J.srows[0][i] = (unsigned char)i; J.mods[0][i] = (unsigned char)i; }
        run_ntt(d, coef, J, true, (long)nq * nb, times);      // times != null: the polynomial is d * times, formed inside the transform
    }
    // fast basis conversion of every digit to the other moduli of Q_level u P: one launch, z = digit.  The digit's own
    // limbs are never copied: the inner product reads them from the NTT-domain input (Decomp::own).
    NttJob J;
    memset(&J, 0, sizeof(J));
    J.nz = beta;
    J.szs = J.dzs = (size_t)rows * n;
    J.nb = nb;
    J.sbs = J.dbs = eb;
    long modup_limbs = 0;
    for (int j = 0; j < beta; j++) {
        const BaseConvTable& T = modup_table(level, j);
        J.n = std::max(J.n, T.nt);
        J.cnt[j] = (unsigned char)T.nt;
        for (int t = 0; t < T.nt; t++) {
            J.rows[j][t] = J.srows[j][t] = T.orow[t];
            J.mods[j][t] = T.tgt[t];
        }
        modup_limbs += T.nt;
    }
    {
        // all digits but the last have alpha sources; the last may have fewer (its own template instance)
        const BaseConvTable* tabs = modup_tables_dev(level);
        const int ns_last = nq - (beta - 1) * prm.alpha;
        const int nfull = ns_last == prm.alpha ? beta : beta - 1;
        // FP64 path: its branch-free form allows a wide (60/61-bit) source only in position 0 of a table
        int mode = bc_mode;
        if (mode == BC_FP)
            for (int j = 0; j < beta; j++)
                if (modup_table(level, j).swide & ~1u) mode = BC_FP_GENERIC;
        if (nfull)
            launch_base_convert(ks, D.ext, coef, tabs, 1, prm.alpha, rows, nfull, 0, (size_t)rows * n, st, mode, nb,
                                (size_t)nq * n, eb);
        if (nfull < beta)
            launch_base_convert(ks, D.ext + (size_t)nfull * rows * n, coef, tabs + nfull, 1, ns_last, rows, 1, 0,
                                (size_t)rows * n, st, mode, nb, (size_t)nq * n, eb);
    }
    // one batched forward NTT over the converted rows of all digits (z = digit) of all batch items
    run_ntt(D.ext, D.ext, J, false, modup_limbs * nb);
    release(coef);
    return D;
}

// <digits, evk> (+ P * addend) into acc = [nb][2][level+1+K][N] over Q_level u P; with accumulate the result is added to
// what acc already holds (several key switches sharing ONE ModDown)
void Engine::ks_inner(const Decomp& D, const EvalKey* evk, const u32* perm, u64* acc, const u64* addend, bool accumulate,
                      bool tensor, size_t addend_bs, int addend_mode) {
    const int level = D.level, nq = level + 1, rows = nq + K();
    std::vector<int> qp = mods_qp(level);
    LimbList ll = limb_list(qp);
    LimbList er = limb_list(qp);                 // evk rows are indexed by global modulus index
    KsBatch kb;
    kb.nb = D.nb;
    kb.acc = (size_t)2 * rows * N();
    kb.ext = (size_t)D.beta * rows * N();
    kb.own = D.own_bs;
    kb.addend = addend_bs;
    kb.addend_mode = addend_mode;
    launch_ks_inner(ks, acc, D.ext, D.own, evk->d, perm, ll, er, D.beta, nmod(), nq, prm.alpha, addend, sl_pmodq,
                    accumulate ? 1 : 0, st, tensor, kb);
    n_keyswitch += D.nb;
}

// ONE division of acc by P * q_{level-drop+1..level}: out is [nb][2][level+1-drop][N] (acc is used as scratch)
void Engine::ks_moddown(u64* acc, int level, int drop, u64* out, int nb, int factor) {
    const size_t n = N();
    const int nq = level + 1, rows = nq + K(), nout = nq - drop;
    if (nout < 1) throw LevelError("key switch: ciphertext level should be positive for this rescale");
    // coefficient form of the limbs that are divided out
    std::vector<int> prow, pmod;
    for (int i = level - drop + 1; i <= level; i++) { prow.push_back(i); pmod.push_back(i); }
    for (int k = 0; k < K(); k++) { prow.push_back(nq + k); pmod.push_back(L() + 1 + k); }
    ntt_rows(acc, prow, pmod, true, 2, (size_t)rows * n, nb, (size_t)2 * rows * n);
    u64* conv = alloc((size_t)nb * 2 * nout * n);
    // [nb][2] slices with uniform strides: the batch folds into the slice count
    const int bcm = (bc_mode == BC_FP && (moddown_table(level, drop).swide & ~1u)) ? BC_FP_GENERIC : bc_mode;
    launch_base_convert(ks, conv, acc, moddown_table_dev(level, drop), 0, K() + drop, nout, 2 * nb, (size_t)rows * n,
                        (size_t)nout * n, st, bcm);
    std::vector<int> qi = mods_q(level - drop);
    if (fuse_ntt) {
        // out = (acc - NTT(conv)) * (P q_dropped)^-1 as the epilogue of the transform's second pass
        NttJob J;
        memset(&J, 0, sizeof(J));
        J.n = nout; J.nz = 2;
        J.szs = J.dzs = (size_t)nout * n;
        J.nb = nb;
        J.sbs = J.dbs = (size_t)2 * nout * n;
        for (int z = 0; z < 2; z++)
            for (int i = 0; i < nout; i++) { J.rows[z][i] = J.srows[z][i] = (unsigned char)i; J.mods[z][i] = (unsigned char)i; }
        NttFuse F;
        F.pro_mod = -1;
        F.ep_a = acc; F.ep_azs = (size_t)rows * n; F.ep_abs = (size_t)2 * rows * n;
        F.ep_out = out; F.ep_ozs = (size_t)nout * n; F.ep_obs = (size_t)2 * nout * n;
        F.s = moddown_inv(level, drop, factor);
        run_ntt_fused(conv, conv, J, F, 2L * nout * nb);
    } else {
        ntt_rows(conv, qi, qi, false, 2, (size_t)nout * n, nb, (size_t)2 * nout * n);
        launch_sub_mul_scalar(ks, out, acc, conv, limb_list(qi), moddown_inv(level, drop, factor), 2 * nb,
                              PolyStride{(size_t)nout * n, (size_t)rows * n, (size_t)nout * n}, st);
    }
    release(conv);
    if (drop) n_rescale += nb;
}

// inner product with the key, then ONE division by P * q_{level-drop+1..level}: out is [nb][2][level+1-drop][N].
// addend ([2][level+1][N] per item, e.g. the (d0, d1) of a tensor product) is folded in as P * addend before the division.
void Engine::ks_apply(const Decomp& D, const EvalKey* evk, const u32* perm, u64* out, const u64* addend, int drop,
                      bool tensor, size_t addend_bs, int addend_mode, int factor) {
    const int rows = D.level + 1 + K();
    u64* acc = alloc((size_t)D.nb * 2 * rows * N());
    ks_inner(D, evk, perm, acc, addend, false, tensor, addend_bs, addend_mode);
    ks_moddown(acc, D.level, drop, out, D.nb, factor);
    release(acc);
}

void Engine::key_switch(const u64* d, int level, const EvalKey* evk, u64* out, int nb, size_t d_bs, const u64* c0_addend,
                        size_t c0_bs) {
    Decomp D = decompose(d, level, nullptr, nb, d_bs, 0);
    // P * c0_addend enters the Q rows of the inner product: exactly c0_addend after the division by P (P x vanishes on the
    // special limbs and ModDown((P x + w)) = x + ModDown(w)), bit-identical to adding it afterwards
    ks_apply(D, evk, nullptr, out, c0_addend, 0, false, c0_bs, c0_addend ? 1 : 0);
    release(D.ext);
}

// ------------------------------------------------------------------ rescale / level management (spec S6)
void Engine::need_levels(int level, int need, const char* what) const {
    if (level < need)
        throw LevelError(std::string(what) + ": ciphertext level should be positive for multiplication (level " +
                         std::to_string(level) + ", need " + std::to_string(need) + ")");
}

// in: [nb][npoly][level+1][N] -> out: [nb][npoly][level][N]
// in_ps: polynomial stride of `in` (0: level + 1 limbs, contiguous); pre_k != 0: the division of pre_k * in (level alignment)
void Engine::rescale_into(u64* out, const u64* in, int npoly, int level, int nb, size_t in_ps, u64 pre_k) {
    const size_t n = N();
    const int nl = level + 1;
    if (!in_ps) in_ps = (size_t)nl * n;
    if (pre_k && !fuse_ntt) throw std::runtime_error("rescale: the scalar rides on the fused epilogue only");
    if (npoly > NTT_MAX_Z) throw std::runtime_error("rescale: too many polynomials");
    u64* last = alloc((size_t)nb * npoly * n);
    {
        NttJob J;
        memset(&J, 0, sizeof(J));
        J.n = 1; J.nz = npoly;
        J.szs = in_ps;
        J.dzs = n;
        J.nb = nb;
        J.sbs = (size_t)npoly * in_ps;
        J.dbs = (size_t)npoly * n;
        for (int z = 0; z < npoly; z++) { J.srows[z][0] = (unsigned char)level; J.rows[z][0] = 0; J.mods[z][0] = (unsigned char)level; }
        run_ntt(in, last, J, true, (long)npoly * nb);
    }
    if (pre_k) {                                               // k * INTT(x) = INTT(k x): the dropped limb only
        std::vector<int> top(1, level);
        ScalarList sk;
        scalar_list(std::vector<u64>(1, pre_k), top, sk);
        launch_mul_scalar(ks, last, last, limb_list(top), sk, npoly * nb, PolyStride{n, n, 0}, st);
    }
    std::vector<int> lo = mods_q(level - 1);
    LimbList ll = limb_list(lo);
    u64* delta = alloc((size_t)nb * npoly * level * n);
    if (fuse_ntt) {
        // the centred lift of the dropped limb is the prologue of the transform's first pass, (in - NTT(delta)) q_l^-1
        // the epilogue of its second: four launches per rescale instead of six, delta never goes to HBM in full
        NttJob J;
        memset(&J, 0, sizeof(J));
        J.n = level; J.nz = npoly;
        J.szs = n;
        J.dzs = (size_t)level * n;
        J.nb = nb;
        J.sbs = (size_t)npoly * n;
        J.dbs = (size_t)npoly * level * n;
        for (int z = 0; z < npoly; z++)
            for (int i = 0; i < level; i++) { J.srows[z][i] = 0; J.rows[z][i] = (unsigned char)i; J.mods[z][i] = (unsigned char)i; }
        NttFuse F;
        F.pro_mod = level;
        F.ep_a = in; F.ep_azs = in_ps; F.ep_abs = (size_t)npoly * in_ps;
        F.ep_out = out; F.ep_ozs = (size_t)level * n; F.ep_obs = (size_t)npoly * level * n;
        F.s = sl_qinv[level];
        F.ep_k = pre_k;
        if (pre_k) scalar_list(std::vector<u64>((size_t)level, pre_k), lo, F.kl);
        run_ntt_fused(last, delta, J, F, (long)npoly * level * nb);
    } else {
        // [nb][npoly] slices with uniform strides: the batch folds into the polynomial count
        launch_rescale_delta(ks, delta, last, ll, level, npoly * nb, PolyStride{(size_t)level * n, n, 0}, st);
        ntt_rows(delta, lo, lo, false, npoly, (size_t)level * n, nb, (size_t)npoly * level * n);
        launch_sub_mul_scalar(ks, out, in, delta, ll, sl_qinv[level], npoly * nb,
                              PolyStride{(size_t)level * n, (size_t)nl * n, (size_t)level * n}, st);
    }
    release(last);
    release(delta);
    n_rescale += nb;
}

Ct* Engine::rescale(const Ct* c) {
    need_levels(c->level, 1, "rescale");
    Ct* r = new_ct(c->npoly, c->level - 1, c->nb);
    rescale_into(r->d, c->d, c->npoly, c->level, c->nb);
    return r;
}

Ct* Engine::drop_to(const Ct* a, int level) {
    if (level > a->level) throw std::runtime_error("drop_to: cannot raise a level");
    const size_t n = N();
    Ct* r = new_ct(a->npoly, level, a->nb);
    if (level == a->level)
        dev::d2d(r->d, a->d, (size_t)a->nb * a->npoly * (level + 1) * n * sizeof(u64), st);
    else
        launch_copy(ks, r->d, a->d, level + 1, a->npoly * a->nb,
                    PolyStride{(size_t)(level + 1) * n, (size_t)(a->level + 1) * n, 0}, st);
    return r;
}

Ct* Engine::copy(const Ct* a) { return drop_to(a, a->level); }

// canonical-scale alignment: drop to target+1, multiply by round(S_t q_{t+1} / S_l), rescale
Ct* Engine::level_down(Ct* c, int target) {
    if (target == c->level) return c;
    if (target > c->level || target < 0) throw std::runtime_error("level_down: bad target");
    for (auto& kv : c->lowered)
        if (kv.first == target) {
            // a copy memoised by another lane has not necessarily been computed yet on the device: order this lane
            // after that lane's stream (an event, no host synchronisation).  A copy made before the open capture began
            // is complete (graph_capture_begin synchronises) and its stream may not belong to the capture: no wait.
            if (in_fork() && kv.second->lane != cur_lane && lane_made[kv.second->lane] && kv.second->cap == capture_id)
                dev::stream_wait(st, streams[kv.second->lane]);
            return kv.second;
        }
    Ct* r = lowered_copy(c, target);
    c->lowered.push_back(std::make_pair(target, r));
    return r;
}
// the alignment itself (spec S6), not memoised: the caller owns the result
Ct* Engine::lowered_copy(const Ct* c, int target) {
    if (target >= c->level || target < 0) throw std::runtime_error("level_down: bad target");
    const size_t n = N();
    const int t1 = target + 1;
    std::vector<int> idx = mods_q(t1);
    const double kf = nearbyint(scales[target] * (double)mod[t1] / scales[c->level]);
    const u64 k = (u64)kf;
    if (fuse_ntt && fuse_align && tabs.cluster < 2 && k) {
        // the scalar rides on the division: k on the dropped limb after its inverse transform, k a in the epilogue; the
        // aligned copy is never written at level target + 1 (bit-identical to the two steps below)
        Ct* r = new_ct(c->npoly, target, c->nb);
        rescale_into(r->d, c->d, c->npoly, t1, c->nb, (size_t)(c->level + 1) * n, k);
        return r;
    }
    std::vector<u64> kv(idx.size(), k);
    ScalarList sc;
    scalar_list(kv, idx, sc);
    const int np = c->npoly * c->nb;                        // uniform strides: the batch folds into the polynomial count
    u64* tmp = alloc((size_t)np * (t1 + 1) * n);
    launch_mul_scalar(ks, tmp, c->d, limb_list(idx), sc, np,
                      PolyStride{(size_t)(t1 + 1) * n, (size_t)(c->level + 1) * n, 0}, st);
    Ct* r = new_ct(c->npoly, target, c->nb);
    rescale_into(r->d, tmp, c->npoly, t1, c->nb);
    release(tmp);
    return r;
}

// ------------------------------------------------------------------ homomorphic ops
// r = a (+/-) b on the common polynomials; the surplus polynomials of the longer operand are copied (negated for b in a - b)
static void addsub(Engine& E, bool subtract, Ct* r, const Ct* a, const Ct* b) {
    const int l = r->level, nb = r->nb;
    const size_t ps = (size_t)(l + 1) * E.N();
    const int np = r->npoly, nmin = std::min(a->npoly, b->npoly);
    LimbList ll = E.limb_list(E.mods_q(l));
    const PolyStride S = psb(ps, ps, ps, nb, np * ps, Engine::bstride(a, a->npoly * ps), Engine::bstride(b, b->npoly * ps));
    if (subtract) launch_sub(E.ks, r->d, a->d, b->d, ll, nmin, S, E.st);
    else launch_add(E.ks, r->d, a->d, b->d, ll, nmin, S, E.st);
    if (np > nmin) {
        const Ct* big = a->npoly > b->npoly ? a : b;
        const PolyStride T = psb(ps, ps, 0, nb, np * ps, Engine::bstride(big, big->npoly * ps), 0);
        if (subtract && big == b) launch_neg(E.ks, r->d + nmin * ps, b->d + nmin * ps, ll, np - nmin, T, E.st);
        else launch_copy(E.ks, r->d + nmin * ps, big->d + nmin * ps, l + 1, np - nmin, T, E.st);
    }
}
Ct* Engine::add(Ct* a, Ct* b) {
    const int l = std::min(a->level, b->level);
    a = level_down(a, l);
    b = level_down(b, l);
    Ct* r = new_ct(std::max(a->npoly, b->npoly), l, batch_of(a, b));
    addsub(*this, false, r, a, b);
    return r;
}
Ct* Engine::sub(Ct* a, Ct* b) {
    const int l = std::min(a->level, b->level);
    a = level_down(a, l);
    b = level_down(b, l);
    Ct* r = new_ct(std::max(a->npoly, b->npoly), l, batch_of(a, b));
    addsub(*this, true, r, a, b);
    return r;
}
Ct* Engine::negate(const Ct* a) {
    const size_t ps = (size_t)(a->level + 1) * N();
    Ct* r = new_ct(a->npoly, a->level, a->nb);
    launch_neg(ks, r->d, a->d, limb_list(mods_q(a->level)), a->npoly * a->nb, PolyStride{ps, ps, 0}, st);
    return r;
}

Ct* Engine::mul_norelin(Ct* a, Ct* b) {
    if (a->npoly != 2 || b->npoly != 2) throw PolyCountError("multiply: operands should have 2 polynomials");
    const int l = std::min(a->level, b->level);
    need_levels(l, 1, "multiply");
    a = level_down(a, l);
    b = level_down(b, l);
    const int nb = batch_of(a, b);
    const size_t n = N(), ps = (size_t)(l + 1) * n;
    u64* t = alloc((size_t)nb * 3 * ps);
    launch_tensor(ks, t, a->d, b->d, limb_list(mods_q(l)), psb(0, 0, 0, nb, 3 * ps, bstride(a, 2 * ps), bstride(b, 2 * ps)), st);
    Ct* r = new_ct(3, l - 1, nb);
    rescale_into(r->d, t, 3, l, nb);
    release(t);
    return r;
}

Ct* Engine::relinearize(const Ct* t) {
    if (t->npoly != 3) throw PolyCountError("relinearize: ciphertext should have 3 polynomials");
    if (!has_relin) throw std::runtime_error("relinearize needs a relinearisation key");
    const int l = t->level, nb = t->nb;
    const size_t ps = (size_t)(l + 1) * N();
    Ct* r = new_ct(2, l, nb);
    key_switch(t->d + 2 * ps, l, &relin, r->d, nb, 3 * ps);
    launch_add(ks, r->d, r->d, t->d, limb_list(mods_q(l)), 2, psb(ps, ps, ps, nb, 2 * ps, 2 * ps, 3 * ps), st);
    return r;
}

// factor (1..3): the product times a small integer, folded into the scalar of the one division (2 a b of the Chebyshev
// recurrences without an addition of its own); bit-identical to adding the product to itself
Ct* Engine::mul(Ct* a, Ct* b, int factor) {
    if (factor < 1 || factor > 3) throw std::runtime_error("multiply: factor must be 1, 2 or 3");
    if (factor != 1 && !fuse_mul_factor) {                       // A/B: the product, then factor - 1 additions
        Ct* p = mul(a, b, 1);
        Ct* r = add(p, p);
        if (factor == 3) { Ct* r3 = add(r, p); free_ct(r); r = r3; }
        free_ct(p);
        return r;
    }
    if (a->npoly != 2 || b->npoly != 2) throw PolyCountError("multiply: operands should have 2 polynomials");
    if (!has_relin) throw std::runtime_error("multiply needs a relinearisation key");
    const int l = std::min(a->level, b->level);
    need_levels(l, 1, "multiply");
    a = level_down(a, l);
    b = level_down(b, l);
    const int nb = batch_of(a, b);
    const size_t n = N(), ps = (size_t)(l + 1) * n;
    const size_t abs_ = bstride(a, 2 * ps), bbs = bstride(b, 2 * ps);
    LimbList ll = limb_list(mods_q(l));
    if (fuse_tensor) {
        // the tensor product is never written: d2 = a1 b1 is formed inside the inverse transform of the decomposition
        // and (as the digits' own rows) inside the inner product, which also adds P (a0 b0, a0 b1 + a1 b0)
        Decomp D = decompose(a->d + ps, l, b->d + ps, nb, abs_, bbs);
        D.own = a->d;
        Ct* r = new_ct(2, l - 1, nb);
        ks_apply(D, &relin, nullptr, r->d, b->d, 1, true, bbs, 0, factor);
        release(D.ext);
        n_mul_cc += nb;
        return r;
    }
    u64* t = alloc((size_t)nb * 3 * ps);
    launch_tensor(ks, t, a->d, b->d, ll, psb(0, 0, 0, nb, 3 * ps, abs_, bbs), st);
    // relinearisation and rescale in one division: (<digits(d2), rlk> + P (d0, d1)) / (P q_l)   (spec S6b)
    Decomp D = decompose(t + 2 * ps, l, nullptr, nb, 3 * ps, 0);
    Ct* r = new_ct(2, l - 1, nb);
    ks_apply(D, &relin, nullptr, r->d, t, 1, false, 3 * ps, 0, factor);
    release(D.ext);
    release(t);
    n_mul_cc += nb;
    return r;
}

Ct* Engine::mul_const(const Ct* a, double re, double im) {
    need_levels(a->level, 1, "multiply");
    const int l = a->level, np = a->npoly * a->nb;          // uniform strides: the batch folds into the polynomial count
    const size_t ps = (size_t)(l + 1) * N();
    std::vector<int> idx = mods_q(l);
    ScalarList cp, cm;
    const_residues(re, im, scales[l], idx, cp, cm);
    u64* t = alloc((size_t)np * ps);
    launch_mul_const(ks, t, a->d, limb_list(idx), cp, cm, np, PolyStride{ps, ps, 0}, st);
    Ct* r = new_ct(a->npoly, l - 1, a->nb);
    rescale_into(r->d, t, a->npoly, l, a->nb);
    release(t);
    return r;
}

Ct* Engine::mul_plain(const Ct* a, const Pt* p) {
    need_levels(a->level, 1, "multiply");
    if (p->level != a->level) throw std::runtime_error("mul_plain: plaintext must be encoded at the ciphertext level");
    const int l = a->level, np = a->npoly * a->nb;
    const size_t ps = (size_t)(l + 1) * N();
    u64* t = alloc((size_t)np * ps);
    launch_mul(ks, t, a->d, p->d, limb_list(mods_q(l)), np, PolyStride{ps, ps, 0}, st);
    Ct* r = new_ct(a->npoly, l - 1, a->nb);
    rescale_into(r->d, t, a->npoly, l, a->nb);
    release(t);
    return r;
}

Ct* Engine::mul_i(const Ct* a, int sign) {
    const int l = a->level;
    const size_t ps = (size_t)(l + 1) * N();
    std::vector<int> idx = mods_q(l);
    std::vector<u64> vp(idx.size()), vm(idx.size());
    for (size_t i = 0; i < idx.size(); i++) {
        const u64 q = mod[idx[i]], j = Jroot[idx[i]];
        vp[i] = sign >= 0 ? j : q - j;
        vm[i] = sign >= 0 ? q - j : j;
    }
    ScalarList cp, cm;
    scalar_list(vp, idx, cp);
    scalar_list(vm, idx, cm);
    Ct* r = new_ct(a->npoly, l, a->nb);
    launch_mul_const(ks, r->d, a->d, limb_list(idx), cp, cm, a->npoly * a->nb, PolyStride{ps, ps, 0}, st);
    return r;
}

Ct* Engine::add_const(const Ct* a, double re, double im) {
    const int l = a->level;
    const size_t ps = (size_t)(l + 1) * N();
    std::vector<int> idx = mods_q(l);
    ScalarList cp, cm;
    const_residues(re, im, scales[l], idx, cp, cm);
    Ct* r = new_ct(a->npoly, l, a->nb);
    // one launch: polynomial 0 of every item receives the constant, the others are copied
    launch_add_const(ks, r->d, a->d, limb_list(idx), cp, cm, a->npoly,
                     psb(ps, ps, 0, a->nb, a->npoly * ps, a->npoly * ps, 0), st);
    return r;
}

Ct* Engine::add_plain(const Ct* a, const Pt* p) {
    if (p->level != a->level) throw std::runtime_error("add_plain: plaintext must be encoded at the ciphertext level");
    const int l = a->level;
    const size_t ps = (size_t)(l + 1) * N(), ab = (size_t)a->npoly * ps;
    Ct* r = new_ct(a->npoly, l, a->nb);
    launch_add(ks, r->d, a->d, p->d, limb_list(mods_q(l)), 1, psb(0, 0, 0, a->nb, ab, ab, 0), st);
    launch_copy(ks, r->d + ps, a->d + ps, l + 1, a->npoly - 1, psb(ps, ps, 0, a->nb, ab, ab, 0), st);
    return r;
}

Ct* Engine::apply_galois(const Ct* a, u64 g) {
    if (a->npoly != 2) throw PolyCountError("rotate/conjugate: ciphertext should have 2 polynomials");
    const int l = a->level, nb = a->nb;
    const size_t ps = (size_t)(l + 1) * N();
    EvalKey* key = galois_key(g);
    u64* t = alloc((size_t)nb * 2 * ps);
    automorph(t, a->d, l + 1, 2 * nb, g);
    Ct* r = new_ct(2, l, nb);
    if (fuse_ks_add) key_switch(t + ps, l, key, r->d, nb, 2 * ps, t, 2 * ps);       // sigma(c0) joins inside the inner product
    else {
        key_switch(t + ps, l, key, r->d, nb, 2 * ps);
        launch_add(ks, r->d, r->d, t, limb_list(mods_q(l)), 1, psb(0, 0, 0, nb, 2 * ps, 2 * ps, 2 * ps), st);
    }
    release(t);
    return r;
}
Ct* Engine::rotate(const Ct* a, long steps) {
    const long n = (long)slots();
    if (((steps % n) + n) % n == 0) return copy(a);
    return apply_galois(a, galois_for_rotation(steps));
}
Ct* Engine::conjugate(const Ct* a) { return apply_galois(a, galois_conj()); }

// hoisted rotations: one ModUp of c1, then per step the Galois gather is fused into the inner product
std::vector<Ct*> Engine::rotate_hoisted(const Ct* a, const std::vector<long>& steps) {
    if (a->npoly != 2) throw PolyCountError("rotate: ciphertext should have 2 polynomials");
    const int l = a->level, nb = a->nb;
    const size_t ps = (size_t)(l + 1) * N();
    const long n = (long)slots();
    std::vector<Ct*> out(steps.size(), nullptr);
    std::vector<size_t> work;
    for (size_t i = 0; i < steps.size(); i++) {
        if (((steps[i] % n) + n) % n == 0) out[i] = copy(a);
        else work.push_back(i);
    }
    if (work.empty()) return out;
    LimbList ll = limb_list(mods_q(l));
    Decomp D = decompose(a->d + ps, l, nullptr, nb, 2 * ps, 0);
    for (size_t i : work) { galois_key(galois_for_rotation(steps[i])); galois_perm(galois_for_rotation(steps[i])); }
    // the rotations only read the shared decomposition: one stream lane each (up to 8)
    const int lanes = (int)std::min<size_t>(work.size(), 8);
    if (lanes > 1) fork(lanes);
    try {
        for (size_t w = 0; w < work.size(); w++) {
            if (lanes > 1) set_lane((int)(w % lanes));
            const size_t i = work[w];
            const u64 g = galois_for_rotation(steps[i]);
            Ct* r = new_ct(2, l, nb);
            if (fuse_ks_add) {
                // c0 is gathered through the same permutation as the digits and joins inside the inner product
                ks_apply(D, galois_key(g), galois_perm(g), r->d, a->d, 0, false, 2 * ps, 2);
            } else {
                ks_apply(D, galois_key(g), galois_perm(g), r->d);
                u64* t = alloc((size_t)nb * ps);
                automorph(t, a->d, l + 1, 1, g, psb(0, 0, 0, nb, ps, 2 * ps, 0));
                launch_add(ks, r->d, r->d, t, ll, 1, psb(0, 0, 0, nb, 2 * ps, 2 * ps, ps), st);
                release(t);
            }
            out[i] = r;
        }
    } catch (...) {
        if (lanes > 1) join();
        release(D.ext);
        for (Ct* c : out) if (c) free_ct(c);
        throw;
    }
    if (lanes > 1) join();
    release(D.ext);
    return out;
}

// [a^1 .. a^degree], a^k = a^(k//2) * a^((k+1)//2)  (depth ceil(log2 k))
// `need` (optional, degree flags): only the flagged powers and the intermediates their products use are computed, with
// the same decomposition k = floor(k/2) + ceil(k/2) as the full basis -- so every power that is computed is bit-identical
// to the one the full basis holds; the others come back as null.  (XOR4 only uses odd exponents: 5 products, not 7.)
std::vector<Ct*> Engine::power_basis(Ct* a, int degree, const unsigned char* need) {
    if (degree < 1) throw std::runtime_error("make_power_basis: degree must be positive");
    int depth = 0;
    while ((1 << depth) < degree) depth++;
    need_levels(a->level, depth, "make_power_basis");
    std::vector<unsigned char> want(degree + 1, need ? 0 : 1);
    if (need) {
        for (int k = degree; k >= 1; k--) {
            if (need[k - 1]) want[k] = 1;
            if (want[k] && k > 1) want[k / 2] = want[(k + 1) / 2] = 1;
        }
        want[1] = 1;
    }
    std::vector<Ct*> out(degree, nullptr);
    out[0] = copy(a);
    bool forked = false;
    try {
        // generation g holds the powers 2^(g-1) < k <= 2^g; they only need powers of earlier generations, so the
        // products of one generation are independent: up to 8 of them run on separate stream lanes
        for (int lo = 1; lo < degree; lo *= 2) {
            const int hi = std::min(2 * lo, degree);
            int cnt = 0;
            for (int k = lo + 1; k <= hi; k++) cnt += want[k];
            const int lanes = std::min(cnt, 8);
            if (lanes > 1) { fork(lanes); forked = true; }
            int slot = 0;
            for (int k = lo + 1; k <= hi; k++) {
                if (!want[k]) continue;
                if (lanes > 1) set_lane(slot++ % lanes);
                out[k - 1] = mul(out[k / 2 - 1], out[(k + 1) / 2 - 1]);
            }
            if (lanes > 1) { join(); forked = false; }
        }
    } catch (...) {
        if (forked) join();
        for (Ct* c : out) if (c) free_ct(c);
        throw;
    }
    return out;
}

}  // namespace ckks
