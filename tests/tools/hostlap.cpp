// tests/tools/hostlap.cpp -- development aid only (never shipped, never imported by the product path).
// Host build of the kernel headers + the host cell sets, driving whole upper_reach_evolve laps on the CPU so the
// rasteriser / paving / enumeration logic can be checked against an oracle `reach` dump without a GPU.
//   hostlap <system.sys> <oracle reach dump> init=l0,u0,l1,u1 [mince=1]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <fstream>
#include <sstream>
#include <map>
#include <algorithm>
#include "../../release-1.0_b200/csrc/ab200_evolve.cuh"
#include "../../release-1.0_b200/csrc/ab200_system.h"
#include "../../release-1.0_b200/csrc/ab200_paving.h"

using namespace ab200;

static HostSystem hs;
static KeyLayout K;
static int ACC;
static std::vector<uint64_t> arena(1 << 22);

static void host_lap(const CellSet& frontier, int mince, double T, int events, CellSet& reach, CellSet& evolve, long& ncells) {
    std::vector<CellBox> cells; frontier.enumerate(cells, mince ? frontier.grid.dim * ACC : -1);
    ncells += cells.size();
    const int MAXB = 4096; const int n = frontier.grid.dim;
    std::vector<double> rb(MAXB * 2 * AB_MAXN), fb(MAXB * 2 * AB_MAXN); std::vector<int32_t> rl(MAXB), fl(MAXB);
    std::vector<uint64_t> rkeys, fkeys;
    for (size_t ci = 0; ci < cells.size(); ++ci) {
        Ctx c; ctx_init_host(c, arena.data(), (int)arena.size());
        EvolveParams P; P.lower = 0; P.backward = 0; P.T = T; P.max_events = events; P.ignore_activations = 0; P.max_boxes = MAXB; P.temporal_order = 6; P.ratio = 5.0; P.spacial_sweep = 1e-12; P.spacial_order = 6; P.max_working_sets = 16;
        CellOutput O; O.reach = rb.data(); O.reach_loc = rl.data(); O.n_reach = 0; O.final_ = fb.data(); O.final_loc = fl.data(); O.n_final = 0; O.stride = 2 * AB_MAXN; O.steps = 0; O.enc = 0; O.enc_cap = 0; O.enc_used = 0; O.n_enc = 0;
        evolve_cell(c, hs.desc, 0, cells[ci].b, P, O);
        if (c.status != ST_OK) { printf("cell %zu status %d\n", ci, c.status); exit(1); }
        for (int which = 0; which < 2; ++which) {
            RunningSet rs; running_init(rs, n);
            int cnt = which ? O.n_final : O.n_reach; const double* src = which ? fb.data() : rb.data();
            std::vector<uint64_t>& dst = which ? fkeys : rkeys;
            for (int k = 0; k < cnt; ++k) {
                Ivl b[AB_GRID_MAXN]; for (int i = 0; i < n; ++i) b[i] = ivl(src[k * O.stride + 2 * i], src[k * O.stride + 2 * i + 1]);
                long long lo[AB_GRID_MAXN], hi[AB_GRID_MAXN], idx[AB_GRID_MAXN];
                if (!raster_box(frontier.grid.origin, frontier.grid.lengths, n, ACC, rs, b, lo, hi)) continue;
                for (int i = 0; i < n; ++i) idx[i] = lo[i];
                for (;;) { uint64_t key; if (key_pack(K, 0, idx, &key)) dst.push_back(key); int d = n - 1; while (d >= 0) { if (++idx[d] <= hi[d]) break; idx[d] = lo[d]; --d; } if (d < 0) break; }
            }
        }
    }
    std::sort(rkeys.begin(), rkeys.end()); rkeys.erase(std::unique(rkeys.begin(), rkeys.end()), rkeys.end());
    std::sort(fkeys.begin(), fkeys.end()); fkeys.erase(std::unique(fkeys.begin(), fkeys.end()), fkeys.end());
    reach = CellSet(frontier.grid, K, 0); reach.keys = rkeys;
    evolve = CellSet(frontier.grid, K, 0); evolve.keys = fkeys;
}

int main(int argc, char** argv) {
    std::ifstream sf(argv[1]); std::stringstream ss; ss << sf.rdbuf();
    std::string err = parse_system(ss.str(), hs); if (!err.empty()) { fprintf(stderr, "%s\n", err.c_str()); return 2; }
    std::ifstream df(argv[2]); std::string line;
    double T = 0, lock = 0; int events = 5; GridDesc g; g.dim = hs.desc.loc[0].n;
    std::vector<std::vector<double> > ref_reach, ref_evolve; std::vector<std::vector<double> >* cur = 0;
    while (std::getline(df, line)) {
        std::istringstream in(line); std::string t; in >> t;
        if (t == "config") { std::string k, name, v; in >> name; while (in >> k >> v) { if (k == "T") T = strtod(v.c_str(), 0); if (k == "lock") lock = strtod(v.c_str(), 0); if (k == "events") events = atoi(v.c_str()); if (k == "acc") ACC = atoi(v.c_str()); } }
        else if (t == "grid") { std::string k, name; in >> k >> name; int n; in >> k >> n; LocDesc& L = hs.desc.loc[0];
            in >> k; for (int i = 0; i < n; ++i) { std::string v; in >> v; g.origin[i] = L.grid_origin[i] = strtod(v.c_str(), 0); }
            in >> k; for (int i = 0; i < n; ++i) { std::string v; in >> v; g.lengths[i] = L.grid_length[i] = strtod(v.c_str(), 0); }
            in >> k; for (int i = 0; i < n; ++i) { std::string v; in >> v; }
            in >> k; { std::string v; in >> v; L.max_step = strtod(v.c_str(), 0); }
            in >> k; for (int i = 0; i < n; ++i) { std::string v; in >> v; L.ref_width[i] = strtod(v.c_str(), 0); } }
        else if (t == "reach") cur = &ref_reach;
        else if (t == "evolve") cur = &ref_evolve;
        else if (t == "c" && cur) { std::vector<double> b; std::string v; while (in >> v) b.push_back(strtod(v.c_str(), 0)); cur->push_back(b); }
    }
    double init[10] = {0}, domain[10] = {-3, 3, -3, 3}; int mince = 0;
    for (int a = 3; a < argc; ++a) {
        if (!strncmp(argv[a], "init=", 5)) { std::stringstream s2(argv[a] + 5); std::string tok; int k = 0; while (std::getline(s2, tok, ',')) init[k++] = atof(tok.c_str()); }
        if (!strncmp(argv[a], "mince=", 6)) mince = atoi(argv[a] + 6);
    }
    K.dim = g.dim; K.acc = ACC; K.loc_bits = 0; for (int i = 0; i < g.dim; ++i) { K.bits[i] = 30; K.bias[i] = 1LL << 29; }
    CellSet initial(g, K, 0), restriction(g, K, 0);
    initial.adjoin_outer_approximation(init);
    restriction.adjoin_outer_approximation(domain);
    initial.restrict(restriction);
    printf("initial leaf cells %zu merged %zu\n", initial.leaf_count(), initial.size());
    unsigned steps = (unsigned)(T / lock); double remaining = T - steps * lock; if (steps == 0) { steps = 1; remaining = 0; lock = T; }
    CellSet reach(g, K, 0), evolve(g, K, 0), found, next; long ncells = 0;
    host_lap(initial, mince, lock, events, found, next, ncells); reach.adjoin(found); evolve.adjoin(next);
    printf("lap 1: reach merged %zu leaf %zu, final merged %zu leaf %zu\n", reach.size(), reach.leaf_count(), evolve.size(), evolve.leaf_count());
    if (getenv("HOSTLAP_DUMP")) { std::vector<CellBox> ec1; evolve.enumerate(ec1); for (size_t k = 0; k < ec1.size(); ++k) printf("e1 %a %a %a %a\n", ec1[k].b[0], ec1[k].b[1], ec1[k].b[2], ec1[k].b[3]); }
    for (unsigned i = 1; i < steps; ++i) { host_lap(evolve, 0, lock, events, found, next, ncells); evolve = next; reach.adjoin(found); }
    if (!evolve.empty() && remaining > 0) { host_lap(evolve, 0, remaining, events, found, next, ncells); evolve = next; reach.adjoin(found); }
    reach.restrict(restriction); evolve.restrict(restriction);
    std::vector<CellBox> rc, ec; reach.enumerate(rc); evolve.enumerate(ec);
    printf("cells evolved %ld; reach merged %zu (ref %zu), evolve merged %zu (ref %zu)\n", ncells, rc.size(), ref_reach.size(), ec.size(), ref_evolve.size());
    if (getenv("HOSTLAP_DUMP")) { Root rr = reach.minimal_root(); printf("reach leaf %zu root h=%d c=(%d,%d)\n", reach.leaf_count(), rr.height, rr.coord[0], rr.coord[1]); for (size_t k = 0; k < rc.size(); ++k) printf("m %a %a %a %a\n", rc[k].b[0], rc[k].b[1], rc[k].b[2], rc[k].b[3]); }
    int bad = 0;
    if (rc.size() != ref_reach.size() || ec.size() != ref_evolve.size()) bad = 1;
    for (size_t k = 0; k < rc.size() && k < ref_reach.size(); ++k) for (int j = 0; j < 2 * g.dim; ++j) if (memcmp(&rc[k].b[j], &ref_reach[k][j], 8)) { if (!bad) printf("reach cell %zu differs: %a vs %a\n", k, rc[k].b[j], ref_reach[k][j]); bad = 1; }
    for (size_t k = 0; k < ec.size() && k < ref_evolve.size(); ++k) for (int j = 0; j < 2 * g.dim; ++j) if (memcmp(&ec[k].b[j], &ref_evolve[k][j], 8)) { if (!bad) printf("evolve cell %zu differs\n", k); bad = 1; }
    printf(bad ? "MISMATCH\n" : "cell sets identical (bit-exact boxes)\n");
    return bad;
}
