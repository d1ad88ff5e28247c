// bootstrap.cu -- full-slot CKKS bootstrapping on the engine (DESIGN.md spec S11).
//
//   1. level_down to level 0, ModRaise to level L: t = Delta_0 m + q_0 I   (|I| <= K)
//   2. CoeffToSlot: `cts_groups` diagonal-sparse matrices (merged inverse special-FFT layers, no bit
//      reversal), factor Delta_L / (2 q_0 K) folded in, evaluated baby-step/giant-step with hoisted
//      baby rotations; one conjugation splits the real and imaginary halves
//   3. EvalMod on both halves: degree-d Chebyshev interpolant of cos(2 pi (K x - 1/4) / 2^r) by Chebyshev
//      division (baby steps T_1..T_m, giants T_2m, T_4m, ..), then r double-angle steps
//   4. SlotToCoeff: `stc_groups` matrices (merged forward layers) with q_0 / (2 pi Delta_0) folded in
//
// Replaces desilofhe.Engine.bootstrap as reached from reference engine_context.py:147-162
// (mixcol_final.py:158-163, invmixcolumns_fhe.py:166-168).  The oracle restates the same spec in
// oracle/bootstrap_oracle.py; the two are compared on decrypted slots.
#include <math.h>

#include <algorithm>
#include <complex>
#include <functional>

#include "engine.cuh"

namespace ckks {

typedef std::complex<double> cplx;
typedef std::map<long, std::vector<cplx>> Diags;     // diagonal index (mod n) -> n entries

struct BsgsTerm {
    int i;            // baby index
    Pt* pt;           // diagonal, pre-rotated by the giant step, encoded at the group's level
};
struct BsgsRow {
    long giant;       // left-rotation applied to the inner sum (slots)
    std::vector<BsgsTerm> terms;
};
struct LinearPlan {
    int level = 0;    // ciphertext level this matrix is applied at
    long stride = 1;
    int n1 = 1;
    std::vector<int> babies;     // baby indices in use
    std::vector<BsgsRow> rows;
};
struct BootPlan {
    int out_level = 0;
    int K = 25, degree = 63, r = 3, m = 8;
    std::vector<double> cheb;
    std::vector<LinearPlan> cts, stc;
};

// ------------------------------------------------------------------ plan (host, fp64)
static Diags fft_layer(size_t n, size_t length, bool inverse) {
    const size_t M = 4 * n, lenh = length / 2, lenq = 4 * length, gap = M / lenq;
    std::vector<size_t> rot(n);
    size_t pw = 1;
    for (size_t j = 0; j < n; j++) { rot[j] = pw; pw = pw * 5 % M; }
    std::vector<cplx> d0(n, 0.0), dp(n, 0.0), dm(n, 0.0);
    for (size_t p = 0; p < n; p++) {
        const size_t j = p % length % lenh;
        const bool first = (p % length) < lenh;
        const double ang = 2.0 * M_PI * (double)((rot[j] % lenq) * gap) / (double)M;
        const cplx w(cos(ang), sin(ang));
        if (!inverse) {
            if (first) { d0[p] = 1.0; dp[p] = w; }
            else { d0[p] = -w; dm[p] = 1.0; }
        } else {
            if (first) { d0[p] = 0.5; dp[p] = 0.5; }
            else { d0[p] = -0.5 * std::conj(w); dm[p] = 0.5 * std::conj(w); }
        }
    }
    Diags out;
    out[0] = d0;
    auto put = [&](long d, const std::vector<cplx>& v) {
        auto it = out.find(d);
        if (it == out.end()) out[d] = v;
        else for (size_t p = 0; p < n; p++) it->second[p] += v[p];
    };
    put((long)(lenh % n), dp);
    put((long)((n - lenh) % n), dm);
    return out;
}

// (A B): diag_{a+b}[p] += A_a[p] * B_b[p+a]
static Diags mat_mul(const Diags& A, const Diags& B, size_t n) {
    Diags out;
    for (auto& ka : A)
        for (auto& kb : B) {
            const long d = (ka.first + kb.first) % (long)n;
            auto it = out.find(d);
            if (it == out.end()) it = out.emplace(d, std::vector<cplx>(n, 0.0)).first;
            const size_t a = (size_t)ka.first;
            for (size_t p = 0; p < n; p++) it->second[p] += ka.second[p] * kb.second[(p + a) % n];
        }
    for (auto it = out.begin(); it != out.end();) {
        double mx = 0;
        for (auto& v : it->second) mx = std::max(mx, std::abs(v));
        if (mx == 0) it = out.erase(it); else ++it;
    }
    return out;
}

static std::vector<Diags> dft_plan(size_t n, int groups, bool inverse, double scale) {
    std::vector<size_t> order;
    for (size_t len = 2; len <= n; len <<= 1) order.push_back(len);
    if (inverse) std::reverse(order.begin(), order.end());
    const int nl = (int)order.size();
    const double per = pow(fabs(scale), 1.0 / groups);
    std::vector<Diags> mats;
    int k = 0;
    for (int g = 0; g < groups; g++) {
        const int cnt = nl / groups + (g < nl % groups ? 1 : 0);
        Diags M;
        for (int t = 0; t < cnt; t++) {
            Diags Lm = fft_layer(n, order[k + t], inverse);
            M = t == 0 ? Lm : mat_mul(Lm, M, n);
        }
        k += cnt;
        for (auto& kv : M)
            for (auto& v : kv.second) v *= per;
        mats.push_back(std::move(M));
    }
    if (scale < 0)
        for (auto& kv : mats[0])
            for (auto& v : kv.second) v = -v;
    return mats;
}

static std::vector<double> cheb_interpolate(const std::function<double(double)>& f, int degree) {
    const int n = degree + 1;
    std::vector<double> fx(n), c(n);
    for (int k = 0; k < n; k++) fx[k] = f(cos(M_PI * (k + 0.5) / n));
    for (int j = 0; j < n; j++) {
        double s = 0;
        for (int k = 0; k < n; k++) s += fx[k] * cos(M_PI * j * (k + 0.5) / n);
        c[j] = s * (j == 0 ? 1.0 : 2.0) / n;
    }
    return c;
}

static LinearPlan make_linear(Engine& E, const Diags& D, int level) {
    const long n = (long)E.slots();
    LinearPlan P;
    P.level = level;
    std::vector<long> ds, signedd;
    long g = 0;
    for (auto& kv : D) {
        const long d = kv.first, s = d <= n / 2 ? d : d - n;
        ds.push_back(d);
        signedd.push_back(s);
        if (s) g = std::__gcd(g, std::labs(s));
    }
    P.stride = g ? g : 1;
    long kmin = 0, kmax = 0;
    for (long s : signedd) { kmin = std::min(kmin, s / P.stride); kmax = std::max(kmax, s / P.stride); }
    const long span = kmax - kmin + 1;
    P.n1 = 1;
    while ((long)P.n1 * P.n1 < span) P.n1 *= 2;
    std::map<long, BsgsRow> rows;
    std::vector<double> buf(2 * n);
    for (size_t t = 0; t < ds.size(); t++) {
        const long k = signedd[t] / P.stride;
        long j = k / P.n1, i = k % P.n1;
        if (i < 0) { i += P.n1; j -= 1; }                       // floor division: i in [0, n1)
        const long giant = j * P.n1 * P.stride;
        const std::vector<cplx>& dg = D.at(ds[t]);
        // rotleft(diag, -giant): out[p] = diag[p - giant]
        for (long p = 0; p < n; p++) {
            const cplx v = dg[(size_t)(((p - giant) % n + n) % n)];
            buf[2 * p] = v.real();
            buf[2 * p + 1] = v.imag();
        }
        BsgsRow& R = rows[j];
        R.giant = giant;
        R.terms.push_back(BsgsTerm{(int)i, E.encode(buf.data(), level)});
        if (std::find(P.babies.begin(), P.babies.end(), (int)i) == P.babies.end()) P.babies.push_back((int)i);
    }
    std::sort(P.babies.begin(), P.babies.end());
    for (auto& kv : rows) {
        std::sort(kv.second.terms.begin(), kv.second.terms.end(), [](const BsgsTerm& a, const BsgsTerm& b) { return a.i < b.i; });
        P.rows.push_back(kv.second);
    }
    return P;
}

// ------------------------------------------------------------------ evaluation helpers
struct Arena {                      // temporaries of one bootstrap; everything but the result is freed
    Engine& E;
    std::vector<Ct*> v;
    explicit Arena(Engine& e) : E(e) {}
    Ct* keep(Ct* c) { v.push_back(c); return c; }
    void release_all_but(Ct* keepme) {
        for (Ct* c : v) if (c != keepme) E.free_ct(c);
        v.clear();
    }
    ~Arena() { for (Ct* c : v) E.free_ct(c); }
};

static Ct* apply_linear(Engine& E, Arena& A, Ct* a, const LinearPlan& P) {
    if (a->level != P.level) throw std::runtime_error("bootstrap: linear transform applied at the wrong level");
    // rotleft(v, k) == rotate(ct, -k); the baby rotations share one ModUp
    std::vector<long> steps;
    for (int i : P.babies) steps.push_back(-(long)i * P.stride);
    std::vector<Ct*> rot = E.rotate_hoisted(a, steps);
    std::map<int, Ct*> baby;
    for (size_t t = 0; t < rot.size(); t++) baby[P.babies[t]] = A.keep(rot[t]);
    const int nb = a->nb;                              // batch items: every buffer below is [nb][2][..], uniform strides
    const size_t ps = (size_t)(P.level + 1) * E.N();
    LimbList ll = E.limb_list(E.mods_q(P.level));
    const PolyStride U{ps, ps, ps};                    // whole ciphertext buffers: the batch folds into 2 * nb polynomials
    PolyStride U0{0, 0, 0};                            // polynomial 0 of every batch item
    U0.nb = nb; U0.bout = U0.ba = U0.bb = 2 * ps;
    // Giant steps with ONE ModDown for the whole matrix ("double hoisting"): every rotated inner sum contributes
    //   sigma_g(c0) in Q_l   and   <ModUp(sigma_g(c1)), rtk_g> in Q_l u P,
    // the Q_l u P parts are accumulated and divided by P once; the un-rotated row (giant 0) is added as it is.
    // The rows are independent: they are dealt to up to 4 stream lanes, each with its own pair of accumulators, which
    // are summed after the join.
    const size_t n = E.N();
    const int rows = P.level + 1 + E.K();
    // lanes for the giant rows: 2 (CKKS_BSGS_LANES).  With both nibble planes in one batched bootstrap every launch is several
    // waves of CTAs; measured 1 / 2 / 3 / 4 / 8 lanes: 46.7 / 46.8 / 47.2 / 47.0 / 47.6 ms per pair and round
    static const int max_lanes = getenv("CKKS_BSGS_LANES") ? std::max(1, atoi(getenv("CKKS_BSGS_LANES"))) : 2;
    const int nl = (int)std::min<size_t>(P.rows.size(), (size_t)max_lanes);
    struct LaneAcc { u64* accqp = nullptr; u64* sum = nullptr; u64* tmp = nullptr; u64* rbuf = nullptr; };
    std::vector<LaneAcc> LA(nl);
    for (const BsgsRow& R : P.rows)
        if (R.giant % (long)E.slots()) {
            E.galois_key(E.galois_for_rotation(-R.giant));
            E.galois_perm(E.galois_for_rotation(-R.giant));
        }
    // the inner sums of ALL rows first, on the parent stream: one launch per 8 rows reads every baby rotation once
    // (CKKS_DIAG_ROWS=0: one launch per row, each re-reading the babies it uses)
    const char* sw = getenv("CKKS_DIAG_ROWS");         // read per call: the A/B test flips it between two bootstraps
    const bool fuse_rows = !sw || atoi(sw) != 0;
    std::vector<u64*> inners(P.rows.size(), nullptr);
    const bool pre = fuse_rows && P.babies.size() <= 16;
    if (pre) {
        std::vector<const Ct*> xs;
        for (int i : P.babies) xs.push_back(baby[i]);
        for (size_t r0 = 0; r0 < P.rows.size(); r0 += 8) {
            const size_t r1 = std::min(P.rows.size(), r0 + 8);
            std::vector<u64*> outs;
            std::vector<std::vector<const Pt*>> ps_;
            for (size_t ri = r0; ri < r1; ri++) {
                inners[ri] = E.alloc((size_t)nb * 2 * ps);
                outs.push_back(inners[ri]);
                std::vector<const Pt*> row(P.babies.size(), nullptr);
                for (const BsgsTerm& T : P.rows[ri].terms)
                    row[std::find(P.babies.begin(), P.babies.end(), T.i) - P.babies.begin()] = T.pt;
                ps_.push_back(row);
            }
            E.diag_mac_rows(outs, xs, ps_, P.level, nb);
        }
    }
    if (nl > 1) E.fork(nl);
    try {
        for (size_t ri = 0; ri < P.rows.size(); ri++) {
            const BsgsRow& R = P.rows[ri];
            LaneAcc& S = LA[ri % nl];
            if (nl > 1) E.set_lane((int)(ri % nl));
            // inner = sum_i diag_i (.) baby_i in one fused multiply-accumulate, un-rescaled (scale S_l^2): one rescale
            // per matrix at the end
            u64* inner = pre ? inners[ri] : E.alloc((size_t)nb * 2 * ps);
            for (size_t off = 0; !pre && off < R.terms.size(); off += 16) {
                std::vector<const Ct*> xs;
                std::vector<const Pt*> ps_;
                for (size_t t = off; t < std::min(R.terms.size(), off + 16); t++) {
                    xs.push_back(baby[R.terms[t].i]);
                    ps_.push_back(R.terms[t].pt);
                }
                if (off == 0) E.diag_mac(inner, xs, ps_, P.level, nb);
                else {
                    if (!S.tmp) S.tmp = E.alloc((size_t)nb * 2 * ps);
                    E.diag_mac(S.tmp, xs, ps_, P.level, nb);
                    launch_add(E.ks, inner, inner, S.tmp, ll, 2 * nb, U, E.st);
                }
            }
            if (R.giant % (long)E.slots() == 0) {
                if (!S.sum) { S.sum = inner; inner = nullptr; if (pre) inners[ri] = nullptr; }
                else launch_add(E.ks, S.sum, S.sum, inner, ll, 2 * nb, U, E.st);
            } else {
                const u64 g = E.galois_for_rotation(-R.giant);
                EvalKey* key = E.galois_key(g);
                if (!S.rbuf) S.rbuf = E.alloc((size_t)nb * 2 * ps);
                E.automorph(S.rbuf, inner, P.level + 1, 2 * nb, g);             // (sigma(c0), sigma(c1))
                Decomp D = E.decompose(S.rbuf + ps, P.level, nullptr, nb, 2 * ps, 0);
                // sigma_g(c0) joins the Q_l u P accumulator as P * sigma_g(c0) inside the inner product (exactly sigma_g(c0)
                // after the one division by P); CKKS_KS_ADD_FUSE=0: a separate addition into the Q_l sum
                const u64* c0 = E.fuse_ks_add ? S.rbuf : nullptr;
                const bool first = !S.accqp;
                if (first) S.accqp = E.alloc((size_t)nb * 2 * rows * n);
                E.ks_inner(D, key, nullptr, S.accqp, c0, !first, false, 2 * ps, c0 ? 1 : 0);
                E.release(D.ext);
                if (!c0) {
                    if (!S.sum) {
                        S.sum = E.alloc((size_t)nb * 2 * ps);
                        dev::zero(S.sum, (size_t)nb * 2 * ps * sizeof(u64), E.st);
                    }
                    launch_add(E.ks, S.sum, S.sum, S.rbuf, ll, 1, U0, E.st);
                }
            }
            if (inner && !pre) E.release(inner);
        }
        for (int l = 0; l < nl; l++) {                 // lane scratch goes back to the lane it came from
            if (nl > 1) E.set_lane(l);
            if (LA[l].tmp) E.release(LA[l].tmp);
            if (LA[l].rbuf) E.release(LA[l].rbuf);
        }
    } catch (...) { if (nl > 1) E.join(); for (u64* p : inners) E.release(p); throw; }
    if (nl > 1) E.join();
    for (u64* p : inners) E.release(p);                // born on the parent stream, returned to it (after the join)
    // combine the lanes' accumulators on the parent stream
    u64* accqp = nullptr;
    Ct* sum = A.keep(E.new_ct(2, P.level, nb));
    bool sum_init = false;
    LimbList llqp = E.limb_list(E.mods_qp(P.level));
    const size_t psqp = (size_t)rows * n;
    for (int l = 0; l < nl; l++) {
        if (LA[l].accqp) {
            if (!accqp) accqp = LA[l].accqp;
            else {
                launch_add(E.ks, accqp, accqp, LA[l].accqp, llqp, 2 * nb, PolyStride{psqp, psqp, psqp}, E.st);
                E.release(LA[l].accqp);
            }
        }
        if (LA[l].sum) {
            if (!sum_init) { dev::d2d(sum->d, LA[l].sum, (size_t)nb * 2 * ps * sizeof(u64), E.st); sum_init = true; }
            else launch_add(E.ks, sum->d, sum->d, LA[l].sum, ll, 2 * nb, U, E.st);
            E.release(LA[l].sum);
        }
    }
    if (accqp) {
        u64* down = E.alloc((size_t)nb * 2 * ps);
        E.ks_moddown(accqp, P.level, 0, down, nb);
        if (sum_init) launch_add(E.ks, sum->d, sum->d, down, ll, 2 * nb, U, E.st);
        else dev::d2d(sum->d, down, (size_t)nb * 2 * ps * sizeof(u64), E.st);
        E.release(down);
        E.release(accqp);
    }
    return A.keep(E.rescale(sum));
}

static Ct* cheb_eval(Engine& E, Arena& A, Ct* x, const std::vector<double>& coef, int m) {
    std::map<int, Ct*> T;
    T[1] = x;
    auto make = [&](int k) -> Ct* {
        const int a = (k + 1) / 2, b = k / 2;                    // T_{a+b} = 2 T_a T_b - T_{a-b}
        Ct* two = A.keep(E.mul(T.at(a), T.at(b), 2));            // 2 T_a T_b: the factor rides on the division's scalar
        return a == b ? A.keep(E.add_const(two, -1.0, 0.0)) : A.keep(E.sub(two, T.at(a - b)));
    };
    // baby steps T_2..T_m by generation (2^(g-1) < k <= 2^g only needs earlier generations): one stream lane per product
    for (int lo = 1; lo < m; lo *= 2) {
        const int hi = std::min(2 * lo, m), cnt = hi - lo;
        std::vector<Ct*> made(cnt, nullptr);
        if (cnt > 1) E.fork(cnt);
        try {
            for (int k = lo + 1; k <= hi; k++) {
                if (cnt > 1) E.set_lane(k - lo - 1);
                made[k - lo - 1] = make(k);
            }
        } catch (...) { if (cnt > 1) E.join(); throw; }
        if (cnt > 1) E.join();
        for (int k = lo + 1; k <= hi; k++) T[k] = made[k - lo - 1];
    }
    for (int g = 2 * m; g <= (int)coef.size() - 1; g *= 2) T[g] = make(g);     // giants T_2m, T_4m, ..
    auto get = [&](int k) -> Ct* { return T.at(k); };
    struct Res { Ct* ct; double c0; };
    std::function<Res(const std::vector<double>&)> rec = [&](const std::vector<double>& c) -> Res {
        const int d = (int)c.size() - 1;
        if (d < m) {
            // leaf: sum_k c_k T_k as one fused linear combination (one rescale per distinct level of the T_k)
            std::vector<Ct*> xs;
            std::vector<double> cs;
            for (int k = 1; k <= d; k++) {
                if (fabs(c[k]) < 1e-300) continue;
                xs.push_back(get(k));
                cs.push_back(c[k]);
                cs.push_back(0.0);
            }
            Ct* acc = xs.empty() ? nullptr : A.keep(E.lincomb(xs, cs.data(), (int)xs.size()));
            return Res{acc, c[0]};
        }
        int g = m;
        while (g * 2 <= d) g *= 2;
        std::vector<double> q(d - g + 1, 0.0), r(c.begin(), c.begin() + g);
        q[0] = c[g];
        for (int k = g + 1; k <= d; k++) { q[k - g] = 2 * c[k]; r[2 * g - k] -= c[k]; }
        // the quotient and remainder sub-polynomials are independent: two stream lanes
        Res Q{nullptr, 0.0}, R{nullptr, 0.0};
        E.fork(2);
        try {
            E.set_lane(0);
            Q = rec(q);
            E.set_lane(1);
            R = rec(r);
        } catch (...) { E.join(); throw; }
        E.join();
        Ct* Tg = get(g);
        // (Q + q_0) T_g: the quotient's constant term joins the quotient before the product (one addition on polynomial 0)
        // instead of q_0 T_g as a constant product with its own rescale, alignment and addition (CKKS_CHEB_C0_FOLD=0)
        static const bool fold_c0 = !getenv("CKKS_CHEB_C0_FOLD") || atoi(getenv("CKKS_CHEB_C0_FOLD")) != 0;
        Ct* qct = Q.ct;
        const bool folded = fold_c0 && qct && Q.c0 != 0.0;
        if (folded) qct = A.keep(E.add_const(qct, Q.c0, 0.0));
        Ct* t = qct ? A.keep(E.mul(qct, Tg)) : nullptr;
        if (Q.c0 != 0.0 && !folded) {
            Ct* t2 = A.keep(E.mul_const(Tg, Q.c0, 0.0));
            t = t ? A.keep(E.add(t, t2)) : t2;
        }
        if (R.ct) t = t ? A.keep(E.add(t, R.ct)) : R.ct;
        return Res{t, R.c0};
    };
    Res out = rec(coef);
    if (!out.ct) throw std::runtime_error("bootstrap: empty Chebyshev polynomial");
    return A.keep(E.add_const(out.ct, out.c0, 0.0));
}

static Ct* eval_mod(Engine& E, Arena& A, Ct* x, const BootPlan& B) {
    Ct* y = cheb_eval(E, A, x, B.cheb, B.m);
    for (int i = 0; i < B.r; i++) {
        Ct* two = A.keep(E.mul(y, y, 2));
        y = A.keep(E.add_const(two, -1.0, 0.0));
    }
    return y;
}

// ------------------------------------------------------------------ Engine entry points
void Engine::bootstrap_setup() {
    if (boot) return;
    if (!has_sk || !has_relin) throw std::runtime_error("bootstrap key needs the secret and relinearisation keys");
    std::unique_ptr<BootPlan> B(new BootPlan());
    const BootParams& bp = prm.boot;
    B->K = bp.K; B->degree = bp.cheb_degree; B->r = bp.double_angle;
    B->m = 1;
    while (B->m * B->m < B->degree + 1) B->m *= 2;
    int giants = 0;
    if (B->degree >= B->m) { int g = B->m; giants = 1; while (g * 2 <= B->degree) { g *= 2; giants++; } }
    int lg = 0;
    while ((1 << lg) < B->m) lg++;
    const int depth = bp.cts_groups + (lg + 1 + giants) + B->r + bp.stc_groups;
    B->out_level = L() - depth;
    if (B->out_level < 1) throw std::runtime_error("bootstrap: the modulus chain is too short (need more than " + std::to_string(depth) + " levels)");
    const double q0 = (double)mod[0];
    const int Kb = B->K, rr = B->r;
    B->cheb = cheb_interpolate([=](double x) { return cos(2.0 * M_PI * (Kb * x - 0.25) / (double)(1 << rr)); }, B->degree);
    const size_t n = slots();
    std::vector<Diags> cts = dft_plan(n, bp.cts_groups, true, scales[L()] / (2.0 * q0 * B->K));
    for (int g = 0; g < bp.cts_groups; g++) B->cts.push_back(make_linear(*this, cts[g], L() - g));
    cts.clear();
    const int stc_top = L() - bp.cts_groups - (lg + 1 + giants) - B->r;
    std::vector<Diags> stc = dft_plan(n, bp.stc_groups, false, q0 / (2.0 * M_PI * scales[0]));
    for (int g = 0; g < bp.stc_groups; g++) B->stc.push_back(make_linear(*this, stc[g], stc_top - g));
    // rotation keys: babies and giants of every matrix, plus conjugation
    galois_key(galois_conj());
    for (const std::vector<LinearPlan>* v : {&B->cts, &B->stc})
        for (const LinearPlan& P : *v) {
            for (int i : P.babies)
                if (i) galois_key(galois_for_rotation(-(long)i * P.stride));
            for (const BsgsRow& R : P.rows)
                if (R.giant % (long)n) galois_key(galois_for_rotation(-R.giant));
        }
    boot = std::move(B);
}

int Engine::boot_out_level() const { return boot ? boot->out_level : -1; }

void Engine::bootstrap_teardown() {
    if (!boot) return;
    for (std::vector<LinearPlan>* v : {&boot->cts, &boot->stc})
        for (LinearPlan& P : *v)
            for (BsgsRow& R : P.rows)
                for (BsgsTerm& T : R.terms) free_pt(T.pt);
    boot.reset();
}

Ct* Engine::mod_raise(Ct* a) {
    if (a->npoly != 2) throw PolyCountError("bootstrap: ciphertext should have 2 polynomials");
    Ct* low = level_down(a, 0);
    const size_t n = N();
    const int top = L(), nb = a->nb;
    u64* coef = alloc((size_t)nb * 2 * n);
    {
        NttJob J;
        memset(&J, 0, sizeof(J));
        J.n = 1; J.nz = 2;
        J.szs = n; J.dzs = n;
        J.nb = nb; J.sbs = J.dbs = 2 * n;
        run_ntt(low->d, coef, J, true, 2L * nb);
    }
    Ct* r = new_ct(2, top, nb);
    std::vector<int> idx = mods_q(top);
    launch_center_lift(ks, r->d, coef, limb_list(idx), 0, 2 * nb, PolyStride{(size_t)(top + 1) * n, n, 0}, st);
    ntt_rows(r->d, idx, idx, false, 2, (size_t)(top + 1) * n, nb, (size_t)2 * (top + 1) * n);
    release(coef);
    return r;
}

Ct* Engine::bootstrap(Ct* a) {
    if (!boot) throw std::runtime_error("bootstrap: no bootstrap key (create_bootstrap_key)");
    const BootPlan& B = *boot;
    Arena A(*this);
    Ct* t = A.keep(mod_raise(a));
    for (const LinearPlan& P : B.cts) t = apply_linear(*this, A, t, P);
    Ct* cj = A.keep(conjugate(t));
    Ct* re = A.keep(add(t, cj));
    Ct* im = A.keep(mul_i(A.keep(sub(t, cj)), -1));
    static const bool stack_halves = getenv("CKKS_EVALMOD_STACK") && atoi(getenv("CKKS_EVALMOD_STACK")) != 0;
    if (stack_halves) {
        // the two halves as ONE handle of 2 nb items: one EvalMod whose launches carry both
        const int nb = re->nb;
        Ct* both = A.keep(stack({re, im}));
        both = eval_mod(*this, A, both, B);
        re = A.keep(slice(both, 0, nb));
        im = A.keep(slice(both, nb, nb));
    } else {
        // the two halves are independent: evaluate them on two stream lanes
        fork(2);
        try {
            set_lane(0);
            re = eval_mod(*this, A, re, B);
            set_lane(1);
            im = eval_mod(*this, A, im, B);
        } catch (...) { join(); throw; }
        join();
    }
    t = A.keep(add(re, A.keep(mul_i(im, +1))));
    if (t->level < B.stc[0].level) throw std::runtime_error("bootstrap: level accounting is off");
    if (t->level > B.stc[0].level) t = level_down(t, B.stc[0].level);     // memoised on (and owned by) its parent
    for (const LinearPlan& P : B.stc) t = apply_linear(*this, A, t, P);
    Ct* out = copy(t);
    n_boot += a->nb;
    return out;
}

}  // namespace ckks
