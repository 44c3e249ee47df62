// tests/tools/hostdbg.cpp -- development aid only (never shipped, never imported by the product path).
// Compiles the kernel headers for the HOST (AB_STRIDE=1, fenv rounding) so the per-cell algorithm can be
// single-stepped in a GPU-less container, and compares its reach/final boxes with an oracle dump produced by
// `oracle/_ref/ref_harness <config> cells ...`.
//
//   hostdbg <system.sys> <oracle cells dump> [maxcells] [verbose]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <fstream>
#include <sstream>
#include <map>
#ifndef AB_NO_EMU
#define AB_WARP_EMU 1
#include "warp_emu.h"
#endif
#include "../../release-1.0_b200/csrc/ab200_evolve.cuh"
#include "../../release-1.0_b200/csrc/ab200_system.h"
#include "../../release-1.0_b200/csrc/ab200_paving.h"
#include <algorithm>

using namespace ab200;
#ifdef AB_MULG_STATS
namespace ab200 { inline namespace AB_VARIANT { long long g_mg[64][4] = {{0}}; } }
#endif
#ifdef AB_COMPOSE_STATS
namespace ab200 { inline namespace AB_VARIANT { long long g_cs[8] = {0}; } }
#endif
#ifdef AB_MUL_HIST
long long ab200::g_mul_hist[8][8] = {{0}};
long long ab200::g_calls[16] = {0}, ab200::g_words[16] = {0};
double ab200::g_flops = 0.0;
long long ab200::g_mul_prod[8][8] = {{0}};
#endif
#ifdef AB_COUNT_ACCESS
long long ab200::g_aw_count = 0;
long long ab200::g_sec[16] = {0};
#endif

struct RefCell { std::vector<std::vector<double> > dreach, dfinal; int have_disc = 0; int loc; std::vector<double> box; std::vector<std::vector<double> > reach, fin; std::vector<int> rloc, floc; std::string err; };

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: hostdbg <system.sys> <cells dump> [maxcells] [verbose]\n"); return 2; }
    std::ifstream sf(argv[1]); std::stringstream ss; ss << sf.rdbuf();
    static HostSystem hs;
    std::string err = parse_system(ss.str(), hs);
    if (!err.empty()) { fprintf(stderr, "parse error: %s\n", err.c_str()); return 2; }
    int maxcells = argc > 3 ? atoi(argv[3]) : 1 << 30;
    int verbose = argc > 4 ? atoi(argv[4]) : 0;
    std::ifstream df(argv[2]);
    std::string line;
    double T = 0; int events = 5, ignore_act = 0, chain = 0, lower_sem = 0, backward_dir = 0;
    std::map<std::string, int> locidx;
    for (size_t i = 0; i < hs.loc_names.size(); ++i) locidx[hs.loc_names[i]] = (int)i;
    std::vector<RefCell> cells; std::vector<std::vector<double> >* dcur = 0; int ACC = 0;
    while (std::getline(df, line)) {
        std::istringstream in(line); std::string t; in >> t;
        if (t == "config") { std::string k, name, v; in >> name; while (in >> k >> v) { if (k == "T") T = strtod(v.c_str(), 0); if (k == "events") events = atoi(v.c_str()); if (k == "chain") chain = atoi(v.c_str()); if (k == "acc") ACC = atoi(v.c_str()); } }
        else if (t == "grid") {
            std::string k, name; in >> k >> name; int li = locidx[name]; LocDesc& L = hs.desc.loc[li];
            int n = 0; in >> k >> n;
            in >> k; for (int i = 0; i < n; ++i) { std::string v; in >> v; L.grid_origin[i] = strtod(v.c_str(), 0); }
            in >> k; for (int i = 0; i < n; ++i) { std::string v; in >> v; L.grid_length[i] = strtod(v.c_str(), 0); }
            in >> k; for (int i = 0; i < n; ++i) { std::string v; in >> v; }
            in >> k; { std::string v; in >> v; L.max_step = strtod(v.c_str(), 0); }
            in >> k; for (int i = 0; i < n; ++i) { std::string v; in >> v; L.ref_width[i] = strtod(v.c_str(), 0); }
        }
        else if (t == "frontier") { std::string k; int nf; in >> nf >> k >> ignore_act; if (in >> k) in >> lower_sem; if (in >> k) in >> backward_dir; }
        else if (t == "cell") { RefCell c; int id; std::string k, name; in >> id >> k >> name >> k; c.loc = locidx[name]; std::string v; while (in >> v) c.box.push_back(strtod(v.c_str(), 0)); cells.push_back(c); }
        else if (t == "r" || t == "f") { std::string name, v; in >> name; std::vector<double> b; while (in >> v) b.push_back(strtod(v.c_str(), 0)); if (t == "r") { cells.back().reach.push_back(b); cells.back().rloc.push_back(locidx[name]); } else { cells.back().fin.push_back(b); cells.back().floc.push_back(locidx[name]); } }
        else if (t == "x") { cells.back().err = line; }
        else if (t == "dreach") { dcur = &cells.back().dreach; cells.back().have_disc = 1; }
        else if (t == "dfinal") { dcur = &cells.back().dfinal; }
        else if (t == "union" || t == "ureach" || t == "ufinal") { dcur = 0; }
        else if (t == "c" && dcur) { std::vector<double> b; std::string v; while (in >> v) b.push_back(strtod(v.c_str(), 0)); dcur->push_back(b); }
        else if (t == "config") { }
    }
    printf("T=%a events=%d ignore_activations=%d chain=%d cells=%zu\n", T, events, ignore_act, chain, cells.size());
    const int ARENA = 1 << 20;
    std::vector<uint64_t> arena(ARENA);
    const int MAXB = 4096;
    std::vector<double> reach(MAXB * 2 * AB_MAXN), fin(MAXB * 2 * AB_MAXN);
    std::vector<int32_t> rloc(MAXB), floc(MAXB);
    int bad = 0, done = 0; long total_steps = 0; int peak = 0;
    for (size_t ci = 0; ci < cells.size() && (int)ci < maxcells; ++ci) {
        RefCell& rc = cells[ci];
        EvolveParams P; P.lower = lower_sem; P.backward = backward_dir; P.T = T; P.max_events = chain ? 1 : events; P.ignore_activations = ignore_act; P.max_boxes = MAXB; P.temporal_order = 6;
        P.ratio = 5.0; P.spacial_sweep = 1e-12; P.spacial_order = 6; P.max_working_sets = 16;
        CellOutput O; O.reach = reach.data(); O.reach_loc = rloc.data(); O.n_reach = 0; O.final_ = fin.data(); O.final_loc = floc.data(); O.n_final = 0; O.stride = 2 * AB_MAXN; O.steps = 0; O.enc = 0; O.enc_cap = 0; O.enc_used = 0; O.n_enc = 0;
        Ctx c;
#ifndef AB_NO_EMU
        {   // one fiber per lane; every lane holds the same (warp-uniform) state, lane 0's copy is reported
            static std::vector<uint64_t> smem(AB_SM_WORDS);
            const int lanes = getenv("AB_EMU_W") ? atoi(getenv("AB_EMU_W")) : 32;
            CellOutput O0 = O; Ctx c0;
            wemu::run(lanes, [&](int lane) {
                Ctx cl; ctx_init(cl, arena.data(), ARENA, getenv("AB_EMU_NOSM") ? 0 : smem.data(), binom_host());
                CellOutput Ol = O;
                evolve_cell(cl, hs.desc, rc.loc, rc.box.data(), P, Ol);
                if (lane == 0) { O0 = Ol; c0 = cl; }
            });
            O = O0; c = c0;
        }
#else
        ctx_init_host(c, arena.data(), ARENA);
        evolve_cell(c, hs.desc, rc.loc, rc.box.data(), P, O);
#endif
        total_steps += O.steps; if (c.A.peak > peak) peak = c.A.peak;
        ++done;
        int n = hs.desc.loc[rc.loc].n;
        bool ok = (c.status == ST_OK) && rc.err.empty() && O.n_reach == (int)rc.reach.size() && O.n_final == (int)rc.fin.size();
        int first_bad = -1;
        if (ok) {
            // the reference's list sets are grouped by location (HybridListSet is a map over locations, iterated in
            // name order); within a location the insertion order is kept. Compare location by location.
            for (int which = 0; which < 2 && ok; ++which) {
                const int cnt = which ? O.n_final : O.n_reach;
                const double* mine = which ? fin.data() : reach.data(); const int32_t* ml = which ? floc.data() : rloc.data();
                const std::vector<std::vector<double> >& rb = which ? rc.fin : rc.reach; const std::vector<int>& rlo = which ? rc.floc : rc.rloc;
                for (int l = 0; l < hs.desc.n_locs && ok; ++l) {
                    std::vector<int> a, b;
                    for (int k = 0; k < cnt; ++k) { if (ml[k] == l) a.push_back(k); if (rlo[k] == l) b.push_back(k); }
                    if (a.size() != b.size()) { ok = false; first_bad = which * 100000; break; }
                    for (size_t q = 0; q < a.size() && ok; ++q) for (int j = 0; j < 2 * hs.desc.loc[l].n; ++j)
                        if (memcmp(&mine[a[q] * O.stride + j], &rb[b[q]][j], 8) != 0) { ok = false; first_bad = which * 100000 + a[q]; break; }
                }
            }
        }
        if (rc.have_disc && c.status == ST_OK) {
            const LocDesc& L = hs.desc.loc[rc.loc];
            KeyLayout K; K.dim = n; K.acc = ACC; K.loc_bits = 0; for (int i = 0; i < n; ++i) { K.bits[i] = 30; K.bias[i] = 1LL << 29; }
            GridDesc g; g.dim = n; for (int i = 0; i < n; ++i) { g.origin[i] = L.grid_origin[i]; g.lengths[i] = L.grid_length[i]; }
            for (int which = 0; which < 2; ++which) {
                RunningSet rs; running_init(rs, n);
                int cnt = which ? O.n_final : O.n_reach; const double* src = which ? fin.data() : reach.data();
                std::vector<uint64_t> keys;
                for (int k = 0; k < cnt; ++k) {
                    Ivl b[AB_GRID_MAXN]; for (int i = 0; i < n; ++i) b[i] = ivl(src[k * O.stride + 2 * i], src[k * O.stride + 2 * i + 1]);
                    long long lo[AB_GRID_MAXN], hi[AB_GRID_MAXN], idx[AB_GRID_MAXN];
                    if (!raster_box(g.origin, g.lengths, n, ACC, rs, b, lo, hi)) continue;
                    for (int i = 0; i < n; ++i) idx[i] = lo[i];
                    for (;;) { uint64_t key; if (key_pack(K, 0, idx, &key)) keys.push_back(key); int d = n - 1; while (d >= 0) { if (++idx[d] <= hi[d]) break; idx[d] = lo[d]; --d; } if (d < 0) break; }
                }
                std::sort(keys.begin(), keys.end()); keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
                CellSet cs(g, K, 0); cs.keys = keys;
                std::vector<CellBox> mine; cs.enumerate(mine);
                const std::vector<std::vector<double> >& ref = which ? rc.dfinal : rc.dreach;
                bool same = mine.size() == ref.size();
                for (size_t k = 0; same && k < mine.size(); ++k) for (int j = 0; j < 2 * n; ++j) if (memcmp(&mine[k].b[j], &ref[k][j], 8)) same = false;
                if (!same) { ok = false; printf("cell %zu %s discretisation differs: mine %zu cells (leaf %zu) ref %zu\n", ci, which ? "final" : "reach", mine.size(), keys.size(), ref.size());
                    if (verbose >= 3) { for (size_t k = 0; k < mine.size(); ++k) printf("  mine %a %a %a %a\n", mine[k].b[0], mine[k].b[1], mine[k].b[2], mine[k].b[3]); for (size_t k = 0; k < ref.size(); ++k) printf("  ref  %a %a %a %a\n", ref[k][0], ref[k][1], ref[k][2], ref[k][3]); } }
            }
        }
        if (!ok && verbose >= 2) {
            printf("cell %zu mismatching reach idx:", ci);
            for (int k = 0; k < O.n_reach && k < (int)rc.reach.size(); ++k) { bool e = true; for (int j = 0; j < 2 * n; ++j) if (memcmp(&reach[k * O.stride + j], &rc.reach[k][j], 8) != 0) e = false; if (!e) printf(" %d", k); }
            printf(" | final:");
            for (int k = 0; k < O.n_final && k < (int)rc.fin.size(); ++k) { bool e = true; for (int j = 0; j < 2 * n; ++j) if (memcmp(&fin[k * O.stride + j], &rc.fin[k][j], 8) != 0) e = false; if (!e) printf(" %d", k); }
            printf("\n");
        }
        if (!ok) {
            ++bad;
            if (bad <= 5 || verbose) {
                printf("cell %zu MISMATCH status=%d reach %d vs %zu final %d vs %zu first_bad=%d referr='%s'\n", ci, c.status, O.n_reach, rc.reach.size(), O.n_final, rc.fin.size(), first_bad, rc.err.c_str());
                int k = first_bad >= 100000 ? -1 : (first_bad >= 0 ? first_bad : 0);
                if (k >= 0 && k < O.n_reach && k < (int)rc.reach.size()) {
                    printf("  ours r[%d]:", k); for (int j = 0; j < 2 * n; ++j) printf(" %a", reach[k * O.stride + j]); printf("\n");
                    printf("  ref  r[%d]:", k); for (int j = 0; j < 2 * n; ++j) printf(" %a", rc.reach[k][j]); printf("\n");
                }
                if (first_bad >= 100000) { int q = first_bad - 100000;
                    printf("  ours f[%d]:", q); for (int j = 0; j < 2 * n; ++j) printf(" %a", fin[q * O.stride + j]); printf("\n");
                    printf("  ref  f[%d]:", q); for (int j = 0; j < 2 * n; ++j) printf(" %a", rc.fin[q][j]); printf("\n"); }
            }
        } else if (verbose) printf("cell %zu ok: %d reach, %d final, %d steps\n", ci, O.n_reach, O.n_final, O.steps);
    }
    printf("cells compared %d, mismatching %d, steps %ld, arena peak words %d\n", done, bad, total_steps, peak);
#ifdef AB_COMPOSE_STATS
    printf("per step: compose calls %.1f, x terms %.1f, term muls %.1f (products %.0f), of which single-term-factor %.1f\n", (double)g_cs[0] / total_steps, (double)g_cs[1] / total_steps, (double)g_cs[2] / total_steps, (double)g_cs[3] / total_steps, (double)g_cs[4] / total_steps);
    printf("per step: chain_ru elements %.0f, chain_ru_masked elements %.0f in %.0f calls\n", (double)g_cs[5] / total_steps, (double)g_cs[6] / total_steps, (double)g_cs[7] / total_steps);
#endif
#ifdef AB_MUL_HIST
    { const char* nm[] = {"add2", "add_many(r terms)", "slide(moved)", "copy", "scal(Float)", "scal(Interval)", "neg", "range", "-", "add_many(streams)", "mul batches(outputs)", "mul batches(products)"};
      for (int i = 0; i < 12; ++i) printf("  %-20s %8.1f calls/step %10.1f terms/step\n", nm[i], (double)g_calls[i] / total_steps, (double)g_words[i] / total_steps); }
    printf("algorithmic FP64 operations (SURVEY 8d formulas): %.0f total, %.0f per step, %.0f per cell\n", g_flops, g_flops / total_steps, g_flops / done);
    printf("tm_mul calls by (x.n bucket rows, y.n bucket cols; bucket b = n<=2^b), per step; then products per step\n");
    for (int a = 0; a < 8; ++a) { for (int b = 0; b < 8; ++b) printf(" %7.1f", (double)g_mul_hist[a][b] / total_steps); printf("\n"); }
    for (int a = 0; a < 8; ++a) { for (int b = 0; b < 8; ++b) printf(" %7.0f", (double)g_mul_prod[a][b] / total_steps); printf("\n"); }
#endif
#ifdef AB_COUNT_ACCESS
    printf("arena word accesses: %lld (%.0f per step)\n", g_aw_count, (double)g_aw_count / total_steps);
    const char* names[] = {"tm_mul(incl)", "tm_compose(incl)", "tm_split", "flow_model(incl)", "flow_apply(incl)", "integration_step", "reachability_step"};
    for (int i = 0; i < 7; ++i) printf("  %-22s %.0f per step\n", names[i], (double)g_sec[i] / total_steps);
#endif
#ifdef AB_MULG_STATS
    for (int b = 0; b < 21; ++b) if (g_mg[b][0]) printf("general muls with monomial bound <= 2^%d: %.1f per step, %.0f products per step\n", b, (double)g_mg[b][0] / total_steps, (double)g_mg[b][1] / total_steps);
    for (int a = 0; a < 20; ++a) if (g_mg[40 + a][0]) printf("general muls with %d variables: %.1f per step, %.0f products per step\n", a, (double)g_mg[40 + a][0] / total_steps, (double)g_mg[40 + a][1] / total_steps);
    printf("largest degree bound %lld\n", g_mg[63][0]);
#endif
    return bad ? 1 : 0;
}
