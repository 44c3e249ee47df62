// tests/tools/primtest.cpp -- development aid only (never shipped).
// Runs every warp-cooperative TaylorModel primitive of release-1.0_b200/csrc on random models under the fiber emulator
// (tests/tools/warp_emu.h) with 1 lane and with W lanes and compares the results bit for bit: the one-lane run is the
// plain sequential algorithm, so any difference is a bug of the cooperative formulation (missing barrier, wrong chain
// order, lost term).
//   primtest [seed] [rounds] [W]
#define AB_WARP_EMU 1
#include "warp_emu.h"
#include "../../release-1.0_b200/csrc/ab200_evolve.cuh"
#include <random>
#include <cstring>

using namespace ab200;

struct HostModel { int as; std::vector<Key> k; std::vector<double> v; double e; Acc acc; };

static std::mt19937_64 rng;
static double urand() { return std::uniform_real_distribution<double>(0.0, 1.0)(rng); }

static HostModel random_model(int as, int maxdeg, int nterms, double emag, bool with_const) {
    HostModel m; m.as = as; m.e = emag * urand(); m.acc = acc_default();
    std::vector<Key> keys;
    for (int t = 0; t < nterms * 3 && (int)keys.size() < nterms; ++t) {
        Key k = key_zero(); int d = (int)(urand() * (maxdeg + 1));
        for (int q = 0; q < d; ++q) key_inc(k, (int)(urand() * as) % as, 1);
        if (!with_const && key_deg(k) == 0) continue;
        bool dup = false; for (auto& o : keys) if (key_eq(o, k)) dup = true;
        if (!dup) keys.push_back(k);
    }
    std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) { return key_less(a, b); });
    m.k = keys;
    for (size_t i = 0; i < keys.size(); ++i) { double mag = std::pow(10.0, -14.0 * urand()); m.v.push_back((urand() < 0.5 ? -1 : 1) * mag * (0.5 + urand())); }
    return m;
}
static TM load_model(Ctx& c, const HostModel& m) {
    TM t = tm_begin(c, m.as, m.acc);
    for (size_t i = 0; i < m.k.size(); ++i) tm_append(c, t, m.k[i], m.v[i]);
    t.e = m.e;
    return t;
}
struct Result { std::vector<uint64_t> words; double e; int n; int status; int rnd; std::vector<double> extra; };
static void grab(Ctx& c, const TM& t, Result& r) { r.n = t.n; r.e = t.e; r.words.assign(c.A.base + t.off, c.A.base + t.off + t.n * AB_TW); r.status = c.status; r.rnd = c.rnd; }
static bool same(const Result& a, const Result& b) {
    return a.n == b.n && a.status == b.status && a.rnd == b.rnd && memcmp(&a.e, &b.e, 8) == 0 && a.words == b.words && a.extra.size() == b.extra.size() &&
           (a.extra.empty() || memcmp(a.extra.data(), b.extra.data(), a.extra.size() * 8) == 0);
}

typedef std::function<void(Ctx&, Result&)> Body;
static std::vector<uint64_t> g_arena(1 << 20), g_sm(AB_SM_WORDS);
static int g_arena_limit = 1 << 20;
static Result run_with(int W, bool use_sm, const Body& body) {
    Result out;
    std::fill(g_arena.begin(), g_arena.begin() + 200000, 0xDEADBEEFDEADBEEFull);
    wemu::run(W, [&](int lane) {
        Ctx c; ctx_init(c, g_arena.data(), g_arena_limit, use_sm ? g_sm.data() : 0, binom_host()); ctx_stage_tables(c);
        Result r; body(c, r);
        if (lane == 0) out = r;
    });
    return out;
}
static int g_fail = 0;
static void check(const char* name, int W, const Body& body) {
    if (getenv("AB_VERBOSE")) { printf("check %s\n", name); fflush(stdout); }
    Result a = run_with(1, true, body), b = run_with(W, true, body), d = run_with(W, false, body);
    const bool ok1 = same(a, b), ok2 = same(a, d);
    if (!ok1 || !ok2) {
        ++g_fail;
        const Result& x = ok1 ? d : b;
        printf("FAIL %-22s (%s) n %d vs %d  e %a vs %a  status %d vs %d rnd %d vs %d\n", name, ok1 ? "arena path" : "shared path", a.n, x.n, a.e, x.e, a.status, x.status, a.rnd, x.rnd);
        for (size_t i = 0; i < a.words.size() && i < x.words.size(); ++i) if (a.words[i] != x.words[i]) { printf("   first differing word %zu: %016llx vs %016llx\n", i, (unsigned long long)a.words[i], (unsigned long long)x.words[i]); break; }
        for (size_t i = 0; i < a.extra.size() && i < x.extra.size(); ++i) if (memcmp(&a.extra[i], &x.extra[i], 8)) { printf("   extra %zu: %a vs %a\n", i, a.extra[i], x.extra[i]); break; }
    }
}

int main(int argc, char** argv) {
    const unsigned long seed = argc > 1 ? strtoul(argv[1], 0, 10) : 1;
    const int rounds = argc > 2 ? atoi(argv[2]) : 20;
    const int W = argc > 3 ? atoi(argv[3]) : 32;
    rng.seed(seed);
    for (int round = 0; round < rounds; ++round) {
        const int as = 1 + (int)(urand() * (AB_MAXARGS - 1));
        const int nx = (int)(urand() * urand() * 150), ny = (int)(urand() * urand() * 150);
        const int dx = 1 + (int)(urand() * 6), dy = 1 + (int)(urand() * 6);
        HostModel X = random_model(as, dx, nx, 1e-6, true), Y = random_model(as, dy, ny, 1e-7, true);
        HostModel Z = random_model(as, dy, ny / 2 + 1, 1e-7, true);
        // make some coefficients of Y the exact negatives of X's so that sums cancel
        for (size_t i = 0; i < X.k.size(); ++i) for (size_t j = 0; j < Y.k.size(); ++j) if (key_eq(X.k[i], Y.k[j]) && urand() < 0.2) Y.v[j] = -X.v[i];
        Acc acc = acc_default(); if (urand() < 0.5) { acc.sweep = 1e-12; acc.maxdeg = 6; }
        const double s = (urand() < 0.2) ? 2.0 : (urand() < 0.2 ? 0.5 : (urand() - 0.5) * 7);
        const Ivl si = ivl(s - 0.1 * urand(), s + 0.1 * urand());
        const int kvar = (int)(urand() * as) % as;
        check("copy+slide", W, [&](Ctx& c, Result& r) { TM x = load_model(c, X); TM y = load_model(c, Y); TM z = tm_copy(c, y); tm_slide(c, z, x.off + AB_TW); grab(c, z, r); });
        check("scal(Float)", W, [&](Ctx& c, Result& r) { TM x = load_model(c, X); tm_scal(c, x, s); grab(c, x, r); });
        check("scal(Interval)", W, [&](Ctx& c, Result& r) { TM x = load_model(c, X); tm_scal(c, x, si); grab(c, x, r); });
        check("acc(Float)", W, [&](Ctx& c, Result& r) { TM x = load_model(c, X); tm_acc(c, x, s); tm_acc(c, x, 0.37); grab(c, x, r); });
        check("acc(Interval)", W, [&](Ctx& c, Result& r) { TM x = load_model(c, Z); tm_acc(c, x, si); tm_acc(c, x, ivl(-0.5, 0.5)); grab(c, x, r); });
        check("neg", W, [&](Ctx& c, Result& r) { TM x = load_model(c, X); TM y = tm_neg(c, x); grab(c, y, r); });
        check("add", W, [&](Ctx& c, Result& r) { TM x = load_model(c, X); TM y = load_model(c, Y); TM z = tm_add(c, x, y); grab(c, z, r); });
        check("sub", W, [&](Ctx& c, Result& r) { TM x = load_model(c, X); TM y = load_model(c, Y); TM z = tm_sub(c, x, y); grab(c, z, r); });
        check("add_many", W, [&](Ctx& c, Result& r) { TM x = load_model(c, X); TM t[3]; t[0] = load_model(c, Y); t[1] = load_model(c, Z); t[2] = load_model(c, X); TM z = tm_add_many(c, x, t, 3, acc); grab(c, z, r); });
        check("mul", W, [&](Ctx& c, Result& r) { TM x = load_model(c, X); TM y = load_model(c, Y); TM z = tm_mul(c, x, y, acc); grab(c, z, r); });
        check("mul(1 term x)", W, [&](Ctx& c, Result& r) { HostModel O = X; if (O.k.size() > 1) { O.k.resize(1); O.v.resize(1); } TM x = load_model(c, O); TM y = load_model(c, Y); TM z = tm_mul(c, x, y, acc); grab(c, z, r); });
        check("mul(1 term y)", W, [&](Ctx& c, Result& r) { HostModel O = Z; if (O.k.size() > 1) { O.k.resize(1); O.v.resize(1); } TM x = load_model(c, X); TM y = load_model(c, O); TM z = tm_mul(c, x, y, acc); grab(c, z, r); });
        check("sqr", W, [&](Ctx& c, Result& r) { TM x = load_model(c, X); TM z = tm_mul(c, x, x, acc); grab(c, z, r); });
        check("sweep", W, [&](Ctx& c, Result& r) { TM x = load_model(c, X); tm_sweep(c, x, 1e-7); grab(c, x, r); });
        check("clean", W, [&](Ctx& c, Result& r) { TM x = load_model(c, X); tm_clean(c, x, acc); grab(c, x, r); });
        check("range+norm", W, [&](Ctx& c, Result& r) { TM x = load_model(c, X); Ivl g = tm_range(c, x); double nm = tm_norm(c, x); grab(c, x, r); r.extra = {g.l, g.u, nm}; });
        check("antidiff", W, [&](Ctx& c, Result& r) { TM x = load_model(c, X); tm_antidiff(c, x, kvar); grab(c, x, r); });
        {
            HostModel A = X, B = X; for (auto& v : B.v) v *= (1 + 1e-9 * (urand() - 0.5)); if (B.k.size() > 2) { B.k.erase(B.k.begin() + 1); B.v.erase(B.v.begin() + 1); } A.e = 1e-3; B.e = 2e-3;
            check("intersection", W, [&](Ctx& c, Result& r) { TM x = load_model(c, A); TM y = load_model(c, B); TM z; z.n = 0; z.off = c.A.top; z.e = 0; bool ok = tm_intersection(c, x, y, z); if (!ok) { z.n = 0; z.e = -1; } grab(c, z, r); });
            HostModel C = X; if (!C.v.empty()) C.v[0] += 1.0; C.e = 2e-3;
            check("intersection(disj)", W, [&](Ctx& c, Result& r) { TM x = load_model(c, A); TM y = load_model(c, C); TM z; z.n = 0; z.off = c.A.top; z.e = 0; bool ok = tm_intersection(c, x, y, z); if (!ok) { z.n = 0; z.e = -1; } grab(c, z, r); });
        }
        const bool upper = urand() < 0.5;
        check("split", W, [&](Ctx& c, Result& r) { TM x = load_model(c, X); TM z = tm_split(c, x, kvar, upper); grab(c, z, r); });
        check("unscale+scale", W, [&](Ctx& c, Result& r) { TM x = load_model(c, X); TM z = tm_unscale(c, x, ivl(0.25, 1.75)); TM w = tm_scale_to(c, z, ivl(-0.3, 0.9)); grab(c, w, r); });
        if (as >= 2) check("partial_evaluate", W, [&](Ctx& c, Result& r) { TM x = load_model(c, X); TM z = tm_partial_evaluate_one(c, x, kvar); grab(c, z, r); });
        {
            // compose: two small outer models in 2 variables substituted by models in `as` variables
            HostModel F0 = random_model(2, 3, 6, 1e-6, true), F1 = random_model(2, 4, 8, 1e-6, true);
            HostModel G0 = random_model(as, 2, 10, 1e-7, true), G1 = random_model(as, 2, 12, 1e-7, true);
            for (auto& v : G0.v) v *= 0.3; for (auto& v : G1.v) v *= 0.3;
            check("compose", W, [&](Ctx& c, Result& r) { TM f[2]; f[0] = load_model(c, F0); f[1] = load_model(c, F1); TM g[2]; g[0] = load_model(c, G0); g[1] = load_model(c, G1); g[0].acc = acc; g[1].acc = acc;
                TM o[2]; compose_at(c, f, 2, g, 2, o); grab(c, o[1], r); r.extra = {o[0].e, (double)o[0].n}; });
        }
    }
    printf("primtest: seed %lu rounds %d W %d: %d failures\n", seed, rounds, W, g_fail);
    return g_fail ? 1 : 0;
}
