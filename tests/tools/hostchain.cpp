// tests/tools/hostchain.cpp -- development aid only (never shipped, never imported by the product path).
// Compiles the analyser mirror (ab200_analyser.cu is plain host C++) against a stand-in for the device lap that runs
// evolve_cell on the CPU, so outer_chain_reach / upper_reach_evolve incl. the grid-side jump sets can be compared
// with an oracle `reach` dump without a GPU.
//   hostchain <system.sys> <oracle reach dump> <chain|upper> acc=A domain=l,u,.. init=l,u,.. [T=..] [events=..]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <fstream>
#include <sstream>
#include <algorithm>
#ifdef AB_CHAIN_EMU      // the enclosure hand-over under the 32-lane warp emulator (lane-strided record copies, wsync placement)
#define AB_WARP_EMU 1
#include "warp_emu.h"
#endif
#include "../../release-1.0_b200/csrc/ab200_evolve.cuh"
#include "../../release-1.0_b200/csrc/ab200_system.h"
#include "../../release-1.0_b200/csrc/ab200_paving.h"
#include "../../include/ariadne_b200.h"

using namespace ab200;

struct ab200_cellset { std::vector<uint64_t> keys; };
struct ab200_boxlist { std::vector<uint64_t> offsets; std::vector<double> boxes; std::vector<int32_t> locs; };
struct ab200_enclosures { std::vector<uint64_t> offsets; std::vector<uint64_t> words; std::vector<int32_t> locs; std::vector<uint64_t> unit_offsets; int key_words; };
struct ab200_ctx { HostSystem sys; KeyLayout K; bool have_layout; std::string err; int accuracy; ab200_settings settings; };
extern "C" int ab200_set_settings(ab200_ctx* ctx, const ab200_settings* s) { ctx->settings = *s; return AB200_OK; }
extern "C" int ab200_ctx_get_settings(const ab200_ctx* ctx, ab200_settings* s) { *s = ctx->settings; return AB200_OK; }
extern "C" int ab200_boxlist_view(const ab200_boxlist* bl, const uint64_t** offsets, const double** boxes, const int32_t** locs) {
    *offsets = bl->offsets.data(); *boxes = bl->boxes.data(); *locs = bl->locs.data(); return AB200_OK; }
extern "C" void ab200_boxlist_free(ab200_boxlist* bl) { delete bl; }

const HostSystem* ab200_ctx_system(const ab200_ctx* ctx) { return &ctx->sys; }
// stand-in for the device batch of ab200_capi.cu: the same check, serially on the host
#include "../../release-1.0_b200/csrc/ab200_jump.cuh"
int ab200_ctx_jump_check(ab200_ctx* ctx, int loc, int event, size_t n, const double* boxes, int* tri, int* lie_missing) {
    static std::vector<uint64_t> arena(1u << 22);
    for (size_t i = 0; i < n; ++i) {
        Ctx c; ctx_init_host(c, arena.data(), (int)arena.size());
        int lm = 0;
        tri[i] = jump_check(c, ctx->sys.desc, loc, event, boxes + i * AB200_BOX_STRIDE, &lm);
        if (lm && lie_missing) *lie_missing = 1;
        if (c.status != ST_OK && !lm) return c.status;
    }
    return AB200_OK;
}
const KeyLayout* ab200_ctx_layout(const ab200_ctx* ctx) { return &ctx->K; }
int ab200_ctx_swap_system(ab200_ctx* ctx, HostSystem* other) { std::swap(ctx->sys, *other); return AB200_OK; }
void ab200_ctx_set_error(ab200_ctx* ctx, const char* msg) { ctx->err = msg; }
std::vector<uint64_t>& ab200_cellset_keys(ab200_cellset* cs) { return cs->keys; }
ab200_cellset* ab200_cellset_new() { return new ab200_cellset(); }
extern "C" void ab200_cellset_free(ab200_cellset* cs) { delete cs; }
extern "C" int ab200_set_location_tuning(ab200_ctx* ctx, int32_t loc, double max_step, const double* ref_widths,
                                         const double* grid_origin, const double* grid_lengths) {
    LocDesc& L = ctx->sys.desc.loc[loc];
    L.max_step = max_step;
    for (int i = 0; i < L.n; ++i) { L.ref_width[i] = ref_widths[i]; L.grid_origin[i] = grid_origin[i]; L.grid_length[i] = grid_lengths[i]; }
    return AB200_OK;
}
extern "C" int ab200_set_accuracy(ab200_ctx* ctx, int32_t accuracy);
extern "C" int ab200_set_accuracy_for_domain(ab200_ctx* ctx, int32_t accuracy, const double*) { return ab200_set_accuracy(ctx, accuracy); }
extern "C" int ab200_set_accuracy(ab200_ctx* ctx, int32_t accuracy) {
    const SystemDesc& S = ctx->sys.desc;
    int dim = S.loc[0].n;
    KeyLayout K; K.dim = dim; K.acc = accuracy;
    int lb = 0; while ((1 << lb) < S.n_locs) ++lb;
    K.loc_bits = lb;
    int per = (63 - lb) / dim; if (per > 40) per = 40;
    for (int i = 0; i < dim; ++i) { K.bits[i] = per; K.bias[i] = 1LL << (per - 1); }
    for (int i = dim; i < AB_GRID_MAXN; ++i) { K.bits[i] = 0; K.bias[i] = 0; }
    ctx->K = K; ctx->accuracy = accuracy; ctx->have_layout = true;
    return AB200_OK;
}

static std::vector<uint64_t> arena(1 << 22);
static long g_cells = 0, g_steps = 0; static int g_peak = 0, g_maxboxes = 0, g_maxenc = 0;

extern "C" int ab200_upper_reach_evolve(ab200_ctx* ctx, uint64_t n_cells, const int32_t* loc, const double* boxes, double T,
                                        int32_t max_events, int32_t ignore_activations, int32_t direction,
                                        ab200_cellset** reach, ab200_cellset** evolve, ab200_status* status) {
    static int lapno = 0; long steps0 = g_steps;
    struct LapLog { int no; uint64_t n; long* s; long s0; ~LapLog() { if (getenv("HOSTCHAIN_LAPS")) printf("lap %d cells %llu steps %ld\n", no, (unsigned long long)n, *s - s0); } } laplog = { lapno++, n_cells, &g_steps, steps0 };
    const SystemDesc& S = ctx->sys.desc;
    const int MAXB = 4096;
    std::vector<double> rb(MAXB * 2 * AB_MAXN), fb(MAXB * 2 * AB_MAXN); std::vector<int32_t> rl(MAXB), fl(MAXB);
    std::vector<uint64_t> rkeys, fkeys;
    for (uint64_t ci = 0; ci < n_cells; ++ci) {
        Ctx c; ctx_init_host(c, arena.data(), (int)arena.size());
        EvolveParams P; P.lower = 0; P.backward = direction ? 1 : 0; P.T = T; P.max_events = max_events; P.ignore_activations = ignore_activations; P.max_boxes = MAXB; P.temporal_order = 6; P.ratio = 5.0; P.spacial_sweep = 1e-12; P.spacial_order = 6; P.max_working_sets = 16;
        CellOutput O; O.reach = rb.data(); O.reach_loc = rl.data(); O.n_reach = 0; O.final_ = fb.data(); O.final_loc = fl.data(); O.n_final = 0; O.stride = 2 * AB_MAXN; O.steps = 0; O.enc = 0; O.enc_cap = 0; O.enc_used = 0; O.n_enc = 0;
        evolve_cell(c, S, loc[ci], boxes + ci * AB200_BOX_STRIDE, P, O);
        g_cells++; g_steps += O.steps; if (c.A.peak > g_peak) g_peak = c.A.peak; if (O.n_reach > g_maxboxes) g_maxboxes = O.n_reach;
        if (c.status != ST_OK) { printf("cell %llu status %d\n", (unsigned long long)ci, c.status); if (status) { status->code = c.status; status->cell = (int64_t)ci; } return c.status; }
        for (int which = 0; which < 2; ++which) {
            int cnt = which ? O.n_final : O.n_reach; const double* src = which ? fb.data() : rb.data(); const int32_t* sl = which ? fl.data() : rl.data();
            std::vector<uint64_t>& dst = which ? fkeys : rkeys;
            for (int l = 0; l < S.n_locs; ++l) {
                const LocDesc& L = S.loc[l]; const int n = L.n;
                RunningSet rs; running_init(rs, n);
                for (int k = 0; k < cnt; ++k) {
                    if (sl[k] != l) continue;
                    Ivl b[AB_GRID_MAXN]; for (int i = 0; i < n; ++i) b[i] = ivl(src[k * O.stride + 2 * i], src[k * O.stride + 2 * i + 1]);
                    long long lo[AB_GRID_MAXN], hi[AB_GRID_MAXN], idx[AB_GRID_MAXN];
                    if (!raster_box(L.grid_origin, L.grid_length, n, ctx->K.acc, rs, b, lo, hi)) continue;
                    for (int i = 0; i < n; ++i) idx[i] = lo[i];
                    for (;;) { uint64_t key; if (key_pack(ctx->K, l, idx, &key)) dst.push_back(key); int d = n - 1; while (d >= 0) { if (++idx[d] <= hi[d]) break; idx[d] = lo[d]; --d; } if (d < 0) break; }
                }
            }
        }
    }
    std::sort(rkeys.begin(), rkeys.end()); rkeys.erase(std::unique(rkeys.begin(), rkeys.end()), rkeys.end());
    std::sort(fkeys.begin(), fkeys.end()); fkeys.erase(std::unique(fkeys.begin(), fkeys.end()), fkeys.end());
    *reach = new ab200_cellset(); (*reach)->keys = rkeys;
    *evolve = new ab200_cellset(); (*evolve)->keys = fkeys;
    if (status) memset(status, 0, sizeof(*status));
    return AB200_OK;
}

extern "C" int ab200_reach_evolve_boxes(ab200_ctx* ctx, uint64_t n_cells, const int32_t* loc, const double* boxes, double T,
                                        int32_t max_events, int32_t ignore_activations, int32_t direction,
                                        ab200_boxlist** reach, ab200_boxlist** fin, ab200_status* status) {
    (void)direction;
    const SystemDesc& S = ctx->sys.desc;
    const int MAXB = 4096;
    std::vector<double> rb(MAXB * 2 * AB_MAXN), fb(MAXB * 2 * AB_MAXN); std::vector<int32_t> rl(MAXB), fl(MAXB);
    ab200_boxlist* R = new ab200_boxlist(); ab200_boxlist* F = new ab200_boxlist();
    R->offsets.push_back(0); F->offsets.push_back(0);
    for (uint64_t ci = 0; ci < n_cells; ++ci) {
        Ctx c; ctx_init_host(c, arena.data(), (int)arena.size());
        EvolveParams P; P.lower = ctx->settings.semantics; P.backward = 0; P.T = T; P.max_events = max_events; P.ignore_activations = ignore_activations; P.max_boxes = MAXB; P.temporal_order = 6; P.ratio = 5.0; P.spacial_sweep = 1e-12; P.spacial_order = 6; P.max_working_sets = 16;
        CellOutput O; O.reach = rb.data(); O.reach_loc = rl.data(); O.n_reach = 0; O.final_ = fb.data(); O.final_loc = fl.data(); O.n_final = 0; O.stride = 2 * AB_MAXN; O.steps = 0; O.enc = 0; O.enc_cap = 0; O.enc_used = 0; O.n_enc = 0;
        evolve_cell(c, S, loc[ci], boxes + ci * AB200_BOX_STRIDE, P, O);
        g_cells++; g_steps += O.steps;
        if (c.status != ST_OK) { printf("cell %llu status %d\n", (unsigned long long)ci, c.status); if (status) { status->code = c.status; status->cell = (int64_t)ci; } return c.status; }
        for (int k = 0; k < O.n_reach; ++k) { for (int j = 0; j < AB200_BOX_STRIDE; ++j) R->boxes.push_back(rb[k * O.stride + j]); R->locs.push_back(rl[k]); }
        for (int k = 0; k < O.n_final; ++k) { for (int j = 0; j < AB200_BOX_STRIDE; ++j) F->boxes.push_back(fb[k * O.stride + j]); F->locs.push_back(fl[k]); }
        R->offsets.push_back(R->locs.size()); F->offsets.push_back(F->locs.size());
    }
    *reach = R; *fin = F;
    if (status) memset(status, 0, sizeof(*status));
    return AB200_OK;
}

// stand-ins for the enclosure hand-over of ab200_capi.cu: the same record format, written and read by the same device code
extern "C" int ab200_reach_evolve_enclosures(ab200_ctx* ctx, uint64_t n_units, const int32_t* loc, const double* boxes, const ab200_enclosures* initial,
                                             double T, int32_t max_events, int ignore_activations, int direction, int32_t max_enclosure_words,
                                             ab200_boxlist** reach, ab200_boxlist** fin, ab200_enclosures** fenc, ab200_status* status) {
    (void)direction;
    const SystemDesc& S = ctx->sys.desc;
    const int MAXB = 4096, ENC = max_enclosure_words > 0 ? max_enclosure_words : 16384;
    std::vector<double> rb(MAXB * 2 * AB_MAXN), fb(MAXB * 2 * AB_MAXN); std::vector<int32_t> rl(MAXB), fl(MAXB); std::vector<uint64_t> eb(ENC);
    ab200_boxlist* R = new ab200_boxlist(); ab200_boxlist* F = new ab200_boxlist(); ab200_enclosures* E = new ab200_enclosures();
    R->offsets.push_back(0); F->offsets.push_back(0); E->unit_offsets.push_back(0); E->key_words = AB_KW;
    if (initial) n_units = initial->locs.size();
    std::vector<double> zero(AB200_BOX_STRIDE, 0.0);
    for (uint64_t ci = 0; ci < n_units; ++ci) {
        Ctx c; ctx_init_host(c, arena.data(), (int)arena.size());
        EvolveParams P; P.lower = ctx->settings.semantics; P.backward = 0; P.T = T; P.max_events = max_events; P.ignore_activations = ignore_activations; P.max_boxes = MAXB; P.temporal_order = 6; P.ratio = 5.0; P.spacial_sweep = 1e-12; P.spacial_order = 6; P.max_working_sets = 16;
        CellOutput O; O.reach = rb.data(); O.reach_loc = rl.data(); O.n_reach = 0; O.final_ = fb.data(); O.final_loc = fl.data(); O.n_final = 0; O.stride = 2 * AB_MAXN; O.steps = 0;
        O.enc = eb.data(); O.enc_cap = ENC; O.enc_used = 0; O.n_enc = 0;
        const int l0 = initial ? initial->locs[ci] : loc[ci];
        const double* b0 = initial ? zero.data() : boxes + ci * AB200_BOX_STRIDE;
        const uint64_t* e0 = initial ? initial->words.data() + initial->offsets[ci] : 0;
#ifdef AB_CHAIN_EMU
        {   // one fiber per lane; every lane holds the same (warp-uniform) state, lane 0's copy is reported
            static std::vector<uint64_t> smem(AB_SM_WORDS);
            CellOutput O0 = O; Ctx c0;
            wemu::run(32, [&](int lane) {
                Ctx cl; ctx_init(cl, arena.data(), (int)arena.size(), smem.data(), binom_host());
                CellOutput Ol = O;
                evolve_cell(cl, S, l0, b0, P, Ol, e0);
                if (lane == 0) { O0 = Ol; c0 = cl; }
            });
            O = O0; c = c0;
        }
#else
        evolve_cell(c, S, l0, b0, P, O, e0);
#endif
        g_cells++; g_steps += O.steps; if (c.A.peak > g_peak) g_peak = c.A.peak; if (O.enc_used > g_maxenc) g_maxenc = O.enc_used; if (O.n_reach > g_maxboxes) g_maxboxes = O.n_reach;
        if (c.status != ST_OK) { printf("unit %llu status %d\n", (unsigned long long)ci, c.status); if (status) { status->code = c.status; status->cell = (int64_t)ci; } return c.status; }
        for (int k = 0; k < O.n_reach; ++k) { for (int j = 0; j < AB200_BOX_STRIDE; ++j) R->boxes.push_back(rb[k * O.stride + j]); R->locs.push_back(rl[k]); }
        for (int k = 0; k < O.n_final; ++k) { for (int j = 0; j < AB200_BOX_STRIDE; ++j) F->boxes.push_back(fb[k * O.stride + j]); F->locs.push_back(fl[k]); }
        R->offsets.push_back(R->locs.size()); F->offsets.push_back(F->locs.size());
        for (int o = 0; o < O.enc_used; ) {
            const int len = (int)(eb[o + 1] & 0xFFFFFFFFu);
            E->offsets.push_back(E->words.size()); E->locs.push_back((int32_t)(eb[o] >> 32));
            E->words.insert(E->words.end(), eb.begin() + o, eb.begin() + o + len);
            o += len;
        }
        E->unit_offsets.push_back(E->locs.size());
    }
    *reach = R; *fin = F; *fenc = E;
    if (status) memset(status, 0, sizeof(*status));
    return AB200_OK;
}
extern "C" int ab200_enclosures_view(const ab200_enclosures* e, uint64_t* count, const int32_t** locs, const uint64_t** unit_offsets, uint64_t* n_units) {
    if (count) *count = e->locs.size();
    if (locs) *locs = e->locs.data();
    if (unit_offsets) *unit_offsets = e->unit_offsets.data();
    if (n_units) *n_units = e->unit_offsets.size() - 1;
    return AB200_OK;
}
extern "C" int ab200_enclosures_select(const ab200_enclosures* src, uint64_t n, const uint64_t* idx, ab200_enclosures** out) {
    ab200_enclosures* e = new ab200_enclosures(); e->key_words = src->key_words; e->unit_offsets.assign(1, 0);
    for (uint64_t k = 0; k < n; ++k) {
        const uint64_t* q = src->words.data() + src->offsets[idx[k]];
        e->offsets.push_back(e->words.size()); e->locs.push_back(src->locs[idx[k]]);
        e->words.insert(e->words.end(), q, q + (size_t)(q[1] & 0xFFFFFFFFu));
        e->unit_offsets.push_back(e->locs.size());
    }
    *out = e; return AB200_OK;
}
extern "C" void ab200_enclosures_free(ab200_enclosures* e) { delete e; }

#include "../../release-1.0_b200/csrc/ab200_analyser.cu"

static std::vector<double> csv(const char* s) { std::vector<double> v; std::stringstream ss(s); std::string t; while (std::getline(ss, t, ',')) v.push_back(strtod(t.c_str(), 0)); return v; }

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage\n"); return 2; }
    static ab200_ctx ctx;
    std::ifstream sf(argv[1]); std::stringstream ss; ss << sf.rdbuf();
    std::string err = parse_system(ss.str(), ctx.sys); if (!err.empty()) { fprintf(stderr, "%s\n", err.c_str()); return 2; }
    const bool chain = !strcmp(argv[3], "chain");
    const bool lower = !strcmp(argv[3], "lower");
    const bool bchain = !strcmp(argv[3], "bchain");
    const bool epslower = !strcmp(argv[3], "epslower");
    int prune = 1; const char* subsys = 0;     // subsys=<file>: the system with its interval parameters at their midpoints
    memset(&ctx.settings, 0, sizeof(ctx.settings));
    int acc = 4, events = 5, mince = 0; double T = 1.0, lock = 0; std::vector<double> dom; std::vector<std::vector<double> > inits; std::vector<int32_t> init_locs; int init_loc = 0;
    for (int a = 4; a < argc; ++a) {
        if (!strncmp(argv[a], "acc=", 4)) acc = atoi(argv[a] + 4);
        if (!strncmp(argv[a], "events=", 7)) events = atoi(argv[a] + 7);
        if (!strncmp(argv[a], "mince=", 6)) mince = atoi(argv[a] + 6);
        if (!strncmp(argv[a], "prune=", 6)) prune = atoi(argv[a] + 6);
        if (!strncmp(argv[a], "subsys=", 7)) subsys = argv[a] + 7;
        if (!strncmp(argv[a], "T=", 2)) T = atof(argv[a] + 2);
        if (!strncmp(argv[a], "lock=", 5)) lock = strtod(argv[a] + 5, 0);
        if (!strncmp(argv[a], "domain=", 7)) dom = csv(argv[a] + 7);
        if (!strncmp(argv[a], "init=", 5)) { inits.push_back(csv(argv[a] + 5)); init_locs.push_back(init_loc); }   // loc= precedes its init=
        if (!strncmp(argv[a], "loc=", 4)) init_loc = atoi(argv[a] + 4);
    }
    const int L = ctx.sys.desc.n_locs, n = ctx.sys.desc.loc[0].n;
    std::vector<double> domain_boxes((size_t)L * AB200_BOX_STRIDE, 0.0), init_box(inits.size() * AB200_BOX_STRIDE, 0.0);
    for (int l = 0; l < L; ++l) for (int i = 0; i < 2 * n; ++i) domain_boxes[(size_t)l * AB200_BOX_STRIDE + i] = dom[(dom.size() >= (size_t)(l + 1) * 2 * n ? l * 2 * n : 0) + i];
    for (size_t k = 0; k < inits.size(); ++k) for (int i = 0; i < 2 * n; ++i) init_box[k * AB200_BOX_STRIDE + i] = inits[k][i];
    ab200_analyser* an = 0;
    int rc = ab200_analyser_create(&an, &ctx, domain_boxes.data(), acc);
    if (rc) { printf("create rc=%d %s\n", rc, ctx.err.c_str()); return 1; }
    if (lock > 0) ab200_analyser_set_lock_time(an, lock);
    printf("lock %a\n", an->lock_time);
    ab200_cellset *reach = 0, *evolve = 0; ab200_status st; memset(&st, 0, sizeof(st));
    std::vector<double> eps((size_t)L * AB200_MAX_DIM, 0.0); int laps = 0;
    // a golden taken with a constraint-set initial set carries the set's text in its "cons| " lines (lower / epslower use it)
    std::string cons_text;
    { std::ifstream cf(argv[2]); std::string cl; while (std::getline(cf, cl)) if (!cl.compare(0, 6, "cons| ")) cons_text += cl.substr(6) + "\n"; }
    if (epslower) {
        // the pruning draw restates glibc's rand(): check it against the C library of this machine first
        { GlibcRand g; for (int i = 0; i < 100000; ++i) if (g.next() != rand()) { printf("GlibcRand differs from rand() at draw %d\n", i); return 1; } }
        std::string subtext;
        if (subsys) { std::ifstream f2(subsys); std::stringstream s2; s2 << f2.rdbuf(); subtext = s2.str(); }
        if (!cons_text.empty()) rc = ab200_analyser_epsilon_lower_chain_reach_constraint(an, cons_text.c_str(), prune, subsys ? subtext.c_str() : 0, &reach, eps.data(), &laps, &st);
        else rc = ab200_analyser_epsilon_lower_chain_reach(an, inits.size(), init_locs.data(), init_box.data(), prune, subsys ? subtext.c_str() : 0, &reach, eps.data(), &laps, &st);
    }
    else if (lower && !cons_text.empty()) rc = ab200_analyser_lower_reach_evolve_constraint(an, cons_text.c_str(), T, events, &reach, &evolve, &st);
    else if (lower) rc = ab200_analyser_lower_reach_evolve(an, inits.size(), init_locs.data(), init_box.data(), T, events, &reach, &evolve, &st);
    else if (!cons_text.empty()) {
        // upper_reach_evolve / outer_chain_reach from the constraint-set initial set of the golden
        ab200_cellset* ini = 0;
        rc = ab200_analyser_discretise_constraint(an, cons_text.c_str(), &ini);
        if (!rc) rc = (chain || bchain) ? ab200_analyser_outer_chain_reach_from(an, ini, bchain ? 1 : 0, &reach, &st)
                                        : ab200_analyser_upper_reach_evolve_from(an, ini, T, events, mince, &reach, &evolve, &st);
    }
    else if (chain || bchain) rc = ab200_analyser_outer_chain_reach(an, inits.size(), init_locs.data(), init_box.data(), bchain ? 1 : 0, &reach, &st);
    else rc = ab200_analyser_upper_reach_evolve(an, inits.size(), init_locs.data(), init_box.data(), T, events, mince, &reach, &evolve, &st);
    if (rc) { printf("rc=%d %s\ncells evolved %ld steps %ld arena peak %d\n", rc, ctx.err.c_str(), g_cells, g_steps, g_peak); return 1; }
    printf("cells evolved %ld steps %ld arena peak %d words, max reach boxes per cell %d, max enclosure words per unit %d\n", g_cells, g_steps, g_peak, g_maxboxes, g_maxenc);
    // compare with the oracle dump: lines "reach"/"evolve" followed by "c <loc?> boxes"
    std::ifstream df(argv[2]); std::string line;
    std::vector<std::vector<double> > ref_reach, ref_evolve; std::vector<std::vector<double> >* cur = 0;
    int bad = 0;
    while (std::getline(df, line)) {
        std::istringstream in(line); std::string t; in >> t;
        if (t == "epsilon" && epslower) {
            // "epsilon loc <name> v0 v1 ..": HybridFloatVector of the reference, compared bit for bit
            std::string kw, name; in >> kw >> name; int l = -1;
            for (size_t k = 0; k < ctx.sys.loc_names.size(); ++k) if (ctx.sys.loc_names[k] == name) l = (int)k;
            std::string v; int i = 0;
            while (l >= 0 && in >> v) { double r = strtod(v.c_str(), 0); if (memcmp(&r, &eps[(size_t)l * AB200_MAX_DIM + i], 8)) { printf("epsilon %s[%d] differs: %a vs %a\n", name.c_str(), i, eps[(size_t)l * AB200_MAX_DIM + i], r); bad = 1; } ++i; }
            if (l < 0) { printf("unknown location %s\n", name.c_str()); bad = 1; }
            continue;
        }
        if (t == "reach") cur = &ref_reach; else if (t == "evolve") cur = &ref_evolve;
        else if (t == "c" && cur) { std::vector<double> b; std::string v; while (in >> v) b.push_back(strtod(v.c_str(), 0)); cur->push_back(b); }
    }
    if (epslower) printf("laps %d\n", laps);
    for (int which = 0; which < ((chain || bchain || epslower) ? 1 : 2); ++which) {
        uint64_t cnt = 0; ab200_analyser_enumerate(an, which ? evolve : reach, 0, &cnt, 0, 0);
        std::vector<int32_t> locs(cnt); std::vector<double> boxes(cnt * AB200_BOX_STRIDE);
        ab200_analyser_enumerate(an, which ? evolve : reach, 0, &cnt, locs.data(), boxes.data());
        const std::vector<std::vector<double> >& ref = which ? ref_evolve : ref_reach;
        printf("%s merged cells %llu (ref %zu)\n", which ? "evolve" : "reach", (unsigned long long)cnt, ref.size());
        if (cnt != ref.size()) bad = 1;
        for (size_t k = 0; k < cnt && k < ref.size(); ++k) {
            size_t off = ref[k].size() - 2 * n;     // optional leading location index
            if (off == 1 && (int)ref[k][0] != locs[k]) { if (!bad) printf("cell %zu location differs\n", k); bad = 1; }
            for (int j = 0; j < 2 * n; ++j) if (memcmp(&boxes[k * AB200_BOX_STRIDE + j], &ref[k][off + j], 8)) { if (!bad) printf("%s cell %zu differs: %a vs %a\n", which ? "evolve" : "reach", k, boxes[k * AB200_BOX_STRIDE + j], ref[k][off + j]); bad = 1; }
        }
        if (getenv("HOSTCHAIN_DUMP")) for (size_t k = 0; k < cnt; ++k) { printf("m %d", locs[k]); for (int j = 0; j < 2 * n; ++j) printf(" %a", boxes[k * AB200_BOX_STRIDE + j]); printf("\n"); }
    }
    printf(bad ? "MISMATCH\n" : "cell sets identical (bit-exact boxes)\n");
    return bad;
}
