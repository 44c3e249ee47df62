// tests/tools/hoststep.cpp -- development aid only. Host build of the kernel headers that prints every
// intermediate Taylor model of one step in the same format as oracle/harness/ref_step.cpp, for diffing.
//   hoststep <system.sys> l0 u0 l1 u1 hmax
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <vector>
#define AB_DEBUG_DUMP 1
#include "../../release-1.0_b200/csrc/ab200_evolve.cuh"
#include "../../release-1.0_b200/csrc/ab200_system.h"
using namespace ab200;
static void dump(Ctx& c, const char* name, const TM& t) {
    printf("%s as=%d n=%d e=%a sweep=%g deg=%d\n", name, t.as, t.n, t.e, t.acc.sweep, t.acc.maxdeg);
    for (int i = 0; i < t.n; ++i) { Key k = tm_key(c, t, i); printf("  ["); for (int j = 0; j < t.as; ++j) printf("%d ", key_get(k, j)); printf("] %a\n", tm_val(c, t, i)); }
}
namespace ab200 { inline namespace AB_VARIANT { void ab_debug_dump(Ctx& c, const char* name, int idx, const TM& t) { char b[64]; if (name[0]=='y' && name[1]==0) sprintf(b, "y%d[%d]", idx/10, idx%10); else if (idx < 0) sprintf(b, "%s", name); else sprintf(b, "%s[%d]", name, idx); dump(c, b, t); } } }
static void dumpv(Ctx& c, const char* name, const TM* v, int n) { for (int i = 0; i < n; ++i) { char b[64]; sprintf(b, "%s[%d]", name, i); dump(c, b, v[i]); } }
int main(int argc, char** argv) {
    std::ifstream sf(argv[1]); std::stringstream ss; ss << sf.rdbuf();
    static HostSystem hs; std::string err = parse_system(ss.str(), hs); if (!err.empty()) { fprintf(stderr, "%s\n", err.c_str()); return 2; }
    const LocDesc& L = hs.desc.loc[0]; int n = L.n;
    std::vector<uint64_t> arena(1 << 20);
    Ctx c; ctx_init_host(c, arena.data(), 1 << 20);
    Acc spacial; spacial.sweep = 1e-12; spacial.maxdeg = 6;
    TM set[AB_MAXN];
    for (int i = 0; i < n; ++i) { set[i] = tm_scaling(c, n, i, ivl(strtod(argv[2 + 2 * i], 0), strtod(argv[3 + 2 * i], 0))); set[i].acc = spacial; }
    double hmax = strtod(argv[2 + 2 * n], 0);
    if (argc > 3 + 2 * n) {
        int nsteps = atoi(argv[3 + 2 * n]); double T = strtod(argv[4 + 2 * n], 0);
        static HostSystem& H = hs;
        for (int i = 0; i < n; ++i) H.desc.loc[0].ref_width[i] = strtod(argv[5 + 2 * n + i], 0);
        H.desc.loc[0].max_step = hmax;
        c.A.top = 0;
        EvolveParams P; P.lower = 0; P.backward = 0; P.T = T; P.max_events = getenv("AB_EVENTS") ? atoi(getenv("AB_EVENTS")) : 5; P.ignore_activations = getenv("AB_IGNORE") ? 1 : 0; P.max_boxes = 4096; P.temporal_order = 6; P.ratio = getenv("AB_RATIO") ? atof(getenv("AB_RATIO")) : 1e30; P.spacial_sweep = 1e-12; P.spacial_order = 6; P.max_working_sets = 4;
        static double rb[4096 * 10], fb[4096 * 10]; static int32_t rl[4096], fl[4096];
        CellOutput O; O.reach = rb; O.reach_loc = rl; O.n_reach = 0; O.final_ = fb; O.final_loc = fl; O.n_final = 0; O.stride = 10; O.steps = 0; O.enc = 0; O.enc_cap = 0; O.enc_used = 0; O.n_enc = 0;
        double box[10]; for (int i = 0; i < 2 * n; ++i) box[i] = strtod(argv[2 + i], 0);
        (void)nsteps;
        evolve_cell(c, H.desc, 0, box, P, O);
        return 0;
    }
    dumpv(c, "set", set, n);
    Ivl range[AB_MAXN]; for (int i = 0; i < n; ++i) range[i] = tm_range(c, set[i]);
    printf("range"); for (int i = 0; i < n; ++i) printf(" %a %a", range[i].l, range[i].u); printf("\n");
    Ivl bb[AB_MAXN], dom[AB_MAXN + 1];
    double h = flow_bounds(c, hs.desc.ops, L.dyn, n, range, hmax, bb);
    printf("h %a bb", h); for (int i = 0; i < n; ++i) printf(" %a %a", bb[i].l, bb[i].u); printf("\n");
    TM flow[AB_MAXN], fset[AB_MAXN], reach[AB_MAXN], ev[AB_MAXN];
    flow_model(c, hs.desc.ops, L.dyn, n, range, h, bb, 6, flow, dom);
    dumpv(c, "flow", flow, n);
    flow_apply(c, flow, n, dom, set, n, h, fset);
    dumpv(c, "fset", fset, n);
    TM one = tm_const(c, n, 1.0, spacial), zero = tm_const(c, n, 0.0, spacial);
    reachability_step(c, fset, n, n, zero, one, spacial, reach);
    dumpv(c, "reach", reach, n);
    integration_step(c, fset, n, n, one, spacial, ev);
    dumpv(c, "evolved", ev, n);
    printf("status %d\n", c.status);
    return 0;
}
