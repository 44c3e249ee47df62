// ab200_capi.cu -- C ABI (include/ariadne_b200.h): context, device buffers, the lap pipeline.
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <cmath>
#include <map>
#include <algorithm>
#include <chrono>
#include <cstring>
#include <cstdio>
#include <dlfcn.h>
#include <nccl.h>
#include "../../include/ariadne_b200.h"
#include "ab200_device.h"
#include "ab200_system.h"
#include "ab200_paving.h"

using namespace ab200;

// Builds of the evolution kernel (one translation unit each, inline namespace = variant name). Plain signatures, so
// that this file, compiled against the AB_KW=1 headers, can call the two-word-key build as well.
//   kw1   one-word MultiIndex keys (<= 7 set parameters): systems with up to 2 state variables
//   kw2   two-word keys (<= 15 set parameters) for systems with more than 2 state variables
#define AB_DECL_VARIANT(ns) namespace ab200 { namespace ns { \
int dev_upload_system_v(const void* system_desc, char* errbuf, int errlen); \
int dev_evolve_capacity_v(char* errbuf, int errlen); \
int dev_launch_evolve_v(cudaStream_t st, int n_cells, int slots, const int32_t* cell_loc, const double* cell_box, const void* evolve_params, \
                        uint64_t* arena, int arena_words, const uint32_t* binom, int* next_cell, double* reach, int32_t* reach_loc, int32_t* n_reach, \
                        double* fin, int32_t* fin_loc, int32_t* n_fin, int32_t* status, int32_t* steps, int32_t* peak, double* flops, \
                        const uint64_t* enc_in, const uint64_t* enc_in_off, uint64_t* enc_out, int enc_cap, int32_t* enc_used, \
                        char* errbuf, int errlen); } }
AB_DECL_VARIANT(kw2)
// The one-cell-per-lane build (ab200_kernels_lane.cu): its own launcher signature (no work counter: one lane per cell,
// arenas interleaved per warp of 32 cells).
namespace ab200 { namespace lane {
int dev_upload_system_v(const void* system_desc, char* errbuf, int errlen);
int dev_evolve_capacity_v(char* errbuf, int errlen);
int dev_launch_evolve_v(cudaStream_t st, int n_cells, const int32_t* cell_loc, const double* cell_box, const void* evolve_params,
                        uint64_t* arena, int arena_words, double* reach, int32_t* reach_loc, int32_t* n_reach,
                        double* fin, int32_t* fin_loc, int32_t* n_fin, int32_t* status, int32_t* steps, int32_t* peak,
                        char* errbuf, int errlen); } }
// A batch this large goes to the lane build (one-word keys only). Below it the warp build is faster: a lap of the lane build
// costs the latency of its slowest cell (seconds), whatever the frontier.
#ifndef AB_LANE_THRESHOLD
#define AB_LANE_THRESHOLD 36864
#endif
#define AB_LANE_ARENA_WORDS 12288
struct EvolveBuild {
    const char* name;
    int (*upload)(const void*, char*, int);
    int (*capacity)(char*, int);
    int (*launch)(cudaStream_t, int, int, const int32_t*, const double*, const void*, uint64_t*, int, const uint32_t*, int*, double*, int32_t*, int32_t*,
                  double*, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*, double*, const uint64_t*, const uint64_t*, uint64_t*, int, int32_t*, char*, int);
};
#define AB_BUILD(ns) { #ns, ab200::ns::dev_upload_system_v, ab200::ns::dev_evolve_capacity_v, ab200::ns::dev_launch_evolve_v }
static const EvolveBuild kBuilds[] = { AB_BUILD(kw1), AB_BUILD(kw2) };
static const int kNumBuilds = (int)(sizeof(kBuilds) / sizeof(kBuilds[0]));
static const EvolveBuild* find_build(const char* name) { for (int i = 0; i < kNumBuilds; ++i) if (!strcmp(kBuilds[i].name, name)) return &kBuilds[i]; return 0; }

// The system tapes live in per-device __constant__ memory; the context that uploaded last owns them. A context that finds
// another owner re-uploads before it launches, so several contexts can share a device.
static const void* g_sys_owner[64] = {0};
static unsigned long long g_sys_generation = 0;

static_assert(AB200_MAX_DIM == AB_MAXN, "header / kernel dimension mismatch");

template<class T> struct DevBuf {
    T* p; size_t n;
    DevBuf() : p(0), n(0) { }
    cudaError_t ensure(size_t want) {
        if (want <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = 0; n = 0;
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
        if (e == cudaSuccess) n = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = 0; n = 0; }
};

struct ab200_cellset { std::vector<uint64_t> keys; };
struct ab200_boxlist { std::vector<uint64_t> offsets; std::vector<double> boxes; std::vector<int32_t> locs; };
// Enclosures (set models) in the device's own record format (ab200_evolve.cuh: out_enclosure); record i starts at word offsets[i].
struct ab200_enclosures { std::vector<uint64_t> offsets; std::vector<uint64_t> words; std::vector<int32_t> locs; std::vector<uint64_t> unit_offsets; int key_words; };

struct ab200_ctx {
    int device;
    cudaStream_t stream;
    HostSystem sys;
    bool have_system;
    ab200_settings settings;
    int accuracy;
    KeyLayout K;
    bool have_layout;
    std::string err;
    // device buffers
    DevBuf<int32_t> d_loc, d_nreach, d_nfin, d_status, d_steps, d_peak, d_rloc, d_floc;
    DevBuf<double> d_box, d_reach, d_fin, d_flops;
    DevBuf<uint64_t> d_arena, d_table, d_keys, d_tmp;
    DevBuf<unsigned> d_hist;
    DevBuf<unsigned long long> d_counters;
    DevBuf<uint32_t> d_binom;
    DevBuf<int> d_next;
    DevBuf<double> d_jbox; DevBuf<int32_t> d_jout;      // jump-set check batches (ab200_ctx_jump_check)
    DevBuf<uint64_t> d_enc_in, d_enc_off, d_enc_out; DevBuf<int32_t> d_enc_used;   // enclosure hand-over (ab200_reach_evolve_enclosures)
    unsigned long long sys_generation;     // bumped by every change of the system description or tuning
    // multi-GPU exchange (ab200_comm_init): one process per GPU, NCCL communicator over NVLink
    ncclComm_t comm; int n_ranks, rank;
    DevBuf<uint64_t> d_gather; DevBuf<unsigned long long> d_counts; DevBuf<unsigned> d_pos;
    cudaEvent_t ev0, ev1;
};

#define CTX_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ctx->err = std::string(#x) + ": " + cudaGetErrorString(e_); return AB200_E_CUDA; } } while (0)
extern "C" void ab200_comm_destroy(ab200_ctx* ctx);
#define DEV_CK(x) do { char eb_[256]; eb_[0] = 0; int r_ = (x); if (r_ != 0) { ctx->err = eb_; return AB200_E_CUDA; } } while (0)

extern "C" void ab200_default_settings(ab200_settings* s) {
    s->temporal_order = 6; s->spacial_order = 6; s->sweep_threshold = 1e-12; s->maximum_enclosure_widths_ratio = 5.0;
    s->max_boxes = 192; s->arena_words = 32768; s->hash_log2 = 24; s->batch_cells = 1 << 17; s->ttl_seconds = 0.0; s->semantics = 0; s->max_working_sets = 0;
}

extern "C" int ab200_ctx_create(ab200_ctx** out, int device) {
    if (!out) return AB200_E_ARGUMENT;
    *out = 0;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return AB200_E_CUDA;   // no GPU: fail loudly
    if (cudaSetDevice(device) != cudaSuccess) return AB200_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return AB200_E_CUDA;
    if (prop.major != 10) return AB200_E_CUDA;    // kernels are built for sm_100a only
    ab200_ctx* ctx = new ab200_ctx();
    ctx->device = device; ctx->have_system = false; ctx->have_layout = false; ctx->accuracy = 0; ctx->sys_generation = 0; ctx->comm = 0; ctx->n_ranks = 1; ctx->rank = 0;
    ab200_default_settings(&ctx->settings);
    if (cudaStreamCreate(&ctx->stream) != cudaSuccess) { delete ctx; return AB200_E_CUDA; }
    cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1);
    char eb[256];
    if (dev_set_stack(32 * 1024, eb, sizeof(eb)) != 0) { delete ctx; return AB200_E_CUDA; }
    {   // binomial table for monomial ranks (ab200_tm.cuh key_rank) and the work counter of the persistent grid
        std::vector<uint32_t> bt(AB_BINOM_N * AB_BINOM_K); binom_fill(bt.data());
        if (ctx->d_binom.ensure(bt.size()) != cudaSuccess || ctx->d_next.ensure(4) != cudaSuccess ||
            cudaMemcpy(ctx->d_binom.p, bt.data(), bt.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) { delete ctx; return AB200_E_CUDA; }
    }
    *out = ctx;
    return AB200_OK;
}
extern "C" void ab200_ctx_destroy(ab200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->d_loc.release(); ctx->d_nreach.release(); ctx->d_nfin.release(); ctx->d_status.release(); ctx->d_steps.release(); ctx->d_peak.release();
    ctx->d_rloc.release(); ctx->d_floc.release(); ctx->d_box.release(); ctx->d_reach.release(); ctx->d_fin.release();
    ctx->d_arena.release(); ctx->d_table.release(); ctx->d_keys.release(); ctx->d_tmp.release(); ctx->d_hist.release(); ctx->d_counters.release(); ctx->d_binom.release(); ctx->d_next.release(); ctx->d_jbox.release(); ctx->d_jout.release(); ctx->d_flops.release(); ctx->d_gather.release(); ctx->d_counts.release(); ctx->d_pos.release();
    ab200_comm_destroy(ctx);
    if (ctx->device >= 0 && ctx->device < 64 && g_sys_owner[ctx->device] == ctx) g_sys_owner[ctx->device] = 0;
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}
extern "C" const char* ab200_last_error(const ab200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

// Make this context's tapes and tuning the ones in the device's constant memory.
static int upload_system(ab200_ctx* ctx, bool changed) {
    CTX_CK(cudaSetDevice(ctx->device));
    if (changed) ctx->sys_generation = ++g_sys_generation;
    if (!changed && ctx->device < 64 && g_sys_owner[ctx->device] == ctx) return AB200_OK;
    CTX_CK(cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < kNumBuilds; ++b) DEV_CK(kBuilds[b].upload(&ctx->sys.desc, eb_, sizeof(eb_)));
    DEV_CK(ab200::lane::dev_upload_system_v(&ctx->sys.desc, eb_, sizeof(eb_)));
    if (ctx->device < 64) g_sys_owner[ctx->device] = ctx;
    return AB200_OK;
}
extern "C" int ab200_set_system_text(ab200_ctx* ctx, const char* text) {
    if (!ctx || !text) return AB200_E_ARGUMENT;
    std::string e = parse_system(text, ctx->sys);
    if (!e.empty()) { ctx->err = "system description: " + e; return AB200_E_ARGUMENT; }
    ctx->have_system = true;
    return upload_system(ctx, true);
}
extern "C" int ab200_system_check(const char* text, int32_t* n_locations, int32_t* n_events, int32_t* n_ops, char* err, int32_t errlen) {
    if (err && errlen > 0) err[0] = 0;
    if (!text) return AB200_E_ARGUMENT;
    HostSystem* hs = new HostSystem();
    std::string e = parse_system(text, *hs);
    int rc = AB200_OK;
    if (!e.empty()) { if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", e.c_str()); rc = AB200_E_ARGUMENT; }
    else {
        int ev = 0; for (int l = 0; l < hs->desc.n_locs; ++l) ev += hs->desc.loc[l].n_events;
        if (n_locations) *n_locations = hs->desc.n_locs;
        if (n_events) *n_events = ev;
        if (n_ops) *n_ops = hs->desc.n_ops;
    }
    delete hs;
    return rc;
}
extern "C" int ab200_system_info(const ab200_ctx* ctx, int32_t* n_locations, int32_t* dims) {
    if (!ctx || !ctx->have_system) return AB200_E_ARGUMENT;
    if (n_locations) *n_locations = ctx->sys.desc.n_locs;
    if (dims) for (int i = 0; i < ctx->sys.desc.n_locs; ++i) dims[i] = ctx->sys.desc.loc[i].n;
    return AB200_OK;
}
extern "C" int ab200_location_index(const ab200_ctx* ctx, const char* name) {
    if (!ctx || !ctx->have_system || !name) return -1;
    for (size_t i = 0; i < ctx->sys.loc_names.size(); ++i) if (ctx->sys.loc_names[i] == name) return (int)i;
    return -1;
}
extern "C" int ab200_set_settings(ab200_ctx* ctx, const ab200_settings* s) {
    if (!ctx || !s) return AB200_E_ARGUMENT;
    if (s->max_boxes < 1 || s->arena_words < 256 || s->hash_log2 < 8 || s->hash_log2 > 32 || s->batch_cells < 32) { ctx->err = "bad settings"; return AB200_E_ARGUMENT; }
    ctx->settings = *s;
    ctx->settings.arena_words = (s->arena_words + 15) & ~15;       // whole 128-byte chunks per lane (AB_CHUNK layouts)
    return AB200_OK;
}
extern "C" int ab200_ctx_get_settings(const ab200_ctx* ctx, ab200_settings* s) {
    if (!ctx || !s) return AB200_E_ARGUMENT;
    *s = ctx->settings;
    return AB200_OK;
}
extern "C" int ab200_set_location_tuning(ab200_ctx* ctx, int32_t loc, double max_step, const double* ref_widths,
                                         const double* grid_origin, const double* grid_lengths) {
    if (!ctx || !ctx->have_system || loc < 0 || loc >= ctx->sys.desc.n_locs) return AB200_E_ARGUMENT;
    LocDesc& L = ctx->sys.desc.loc[loc];
    L.max_step = max_step;
    for (int i = 0; i < L.n; ++i) { L.ref_width[i] = ref_widths[i]; L.grid_origin[i] = grid_origin[i]; L.grid_length[i] = grid_lengths[i]; }
    return upload_system(ctx, true);
}
// Key window: equal share of the 64 bits after the location field, centred on the grid origin.
extern "C" int ab200_set_accuracy(ab200_ctx* ctx, int32_t accuracy) {
    if (!ctx || !ctx->have_system || accuracy < 0 || accuracy > 24) return AB200_E_ARGUMENT;
    const SystemDesc& S = ctx->sys.desc;
    int dim = S.loc[0].n;
    for (int l = 1; l < S.n_locs; ++l) if (S.loc[l].n != dim) { ctx->err = "locations of different dimension are not supported"; return AB200_E_UNSUPPORTED; }
    KeyLayout K; K.dim = dim; K.acc = accuracy;
    int lb = 0; while ((1 << lb) < S.n_locs) ++lb;
    K.loc_bits = lb;
    int per = (63 - lb) / dim; if (per > 40) per = 40;
    if (per < accuracy + 3) { ctx->err = "accuracy too fine for 64-bit cell keys"; return AB200_E_KEYSPACE; }
    for (int i = 0; i < dim; ++i) { K.bits[i] = per; K.bias[i] = 1LL << (per - 1); }
    for (int i = dim; i < AB_GRID_MAXN; ++i) { K.bits[i] = 0; K.bias[i] = 0; }
    ctx->K = K; ctx->accuracy = accuracy; ctx->have_layout = true;
    return AB200_OK;
}
// Cell keys sized to a domain (a 5-variable system at accuracy 10 does not fit the uniform split of 63 bits).
// Per coordinate the key field covers the lattice indices of the domain boxes of all locations at this accuracy plus a
// margin of a quarter base cell (at least 16 cells) per side; spare bits widen the margins evenly. Cells that fall outside
// the key space are reported (AB200_E_KEYSPACE), never dropped.
extern "C" int ab200_set_accuracy_for_domain(ab200_ctx* ctx, int32_t accuracy, const double* domain_boxes) {
    if (!ctx || !ctx->have_system || !domain_boxes || accuracy < 0 || accuracy > 24) return AB200_E_ARGUMENT;
    if (ab200_set_accuracy(ctx, accuracy) == AB200_OK) return AB200_OK;          // the uniform layout is kept whenever it fits
    const SystemDesc& S = ctx->sys.desc;
    const int dim = S.loc[0].n;
    for (int l = 1; l < S.n_locs; ++l) if (S.loc[l].n != dim) return AB200_E_UNSUPPORTED;
    KeyLayout K; K.dim = dim; K.acc = accuracy;
    int lb = 0; while ((1 << lb) < S.n_locs) ++lb;
    K.loc_bits = lb;
    const double fine = (double)(1LL << accuracy);
    long long margin = (1LL << accuracy) / 4; if (margin < 16) margin = 16;
    int total = lb;
    long long lo[AB_GRID_MAXN], hi[AB_GRID_MAXN];
    for (int i = 0; i < dim; ++i) {
        bool any = false;
        for (int l = 0; l < S.n_locs; ++l) {
            const double a = domain_boxes[(size_t)l * AB200_BOX_STRIDE + 2 * i], b = domain_boxes[(size_t)l * AB200_BOX_STRIDE + 2 * i + 1];
            if (!(a <= b)) continue;                                                // empty domain in this location
            const double o = S.loc[l].grid_origin[i], len = S.loc[l].grid_length[i];
            if (!(len > 0)) { ctx->err = "ab200_set_location_tuning must be called for every location first"; return AB200_E_ARGUMENT; }
            const double fa = std::floor((a - o) / len * fine), fb = std::ceil((b - o) / len * fine);
            if (!(std::fabs(fa) < 4.0e18) || !(std::fabs(fb) < 4.0e18)) { ctx->err = "domain too large for 64-bit cell keys"; return AB200_E_KEYSPACE; }
            const long long ia = (long long)fa - margin, ib = (long long)fb + margin;
            if (!any || ia < lo[i]) lo[i] = ia;
            if (!any || ib > hi[i]) hi[i] = ib;
            any = true;
        }
        if (!any) { lo[i] = -margin; hi[i] = margin; }
        int b = 1; while (b < 62 && (1LL << b) < hi[i] - lo[i] + 1) ++b;
        K.bits[i] = b; total += b;
    }
    if (total > 63) { ctx->err = "accuracy too fine for 64-bit cell keys on this domain"; return AB200_E_KEYSPACE; }
    for (int i = 0; total < 63 && K.bits[i % dim] < 40; ++i) { if (i >= 64 * dim) break; K.bits[i % dim]++; total++; }      // spare bits: wider margins
    for (int i = 0; i < dim; ++i) { const long long span = 1LL << K.bits[i], used = hi[i] - lo[i] + 1; K.bias[i] = -lo[i] + (span - used) / 2; }
    for (int i = dim; i < AB_GRID_MAXN; ++i) { K.bits[i] = 0; K.bias[i] = 0; }
    ctx->K = K; ctx->accuracy = accuracy; ctx->have_layout = true; ctx->err.clear();
    return AB200_OK;
}
extern "C" int ab200_key_decode(const ab200_ctx* ctx, uint64_t key, int32_t* loc, int64_t* idx) {
    if (!ctx || !ctx->have_layout) return AB200_E_ARGUMENT;
    long long v[AB_GRID_MAXN]; int l;
    key_unpack(ctx->K, key, &l, v);
    if (loc) *loc = l;
    if (idx) for (int i = 0; i < ctx->K.dim; ++i) idx[i] = v[i];
    return AB200_OK;
}

// ---- NCCL, loaded on demand (a single-GPU user never needs it) ----------------------------------------------------------
namespace {
struct NcclApi {
    void* h;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
};
NcclApi* nccl_api(std::string* err) {
    static NcclApi api; static bool tried = false, ok = false;
    if (!tried) {
        tried = true;
        api.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.h) api.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (api.h) {
            api.GetUniqueId = (ncclResult_t (*)(ncclUniqueId*))dlsym(api.h, "ncclGetUniqueId");
            api.CommInitRank = (ncclResult_t (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(api.h, "ncclCommInitRank");
            api.CommDestroy = (ncclResult_t (*)(ncclComm_t))dlsym(api.h, "ncclCommDestroy");
            api.AllGather = (ncclResult_t (*)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t))dlsym(api.h, "ncclAllGather");
            api.GetErrorString = (const char* (*)(ncclResult_t))dlsym(api.h, "ncclGetErrorString");
            ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
        }
    }
    if (!ok) { if (err) *err = "libnccl.so.2 could not be loaded"; return 0; }
    return &api;
}
}
#define NCCL_CK(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { ctx->err = std::string(#x) + ": " + nc->GetErrorString(r_); return AB200_E_CUDA; } } while (0)

extern "C" int ab200_comm_unique_id(uint8_t* id128) {
    if (!id128) return AB200_E_ARGUMENT;
    NcclApi* nc = nccl_api(0);
    if (!nc) return AB200_E_CUDA;
    ncclUniqueId id;
    if (nc->GetUniqueId(&id) != ncclSuccess) return AB200_E_CUDA;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return AB200_OK;
}
extern "C" int ab200_comm_init(ab200_ctx* ctx, int32_t n_ranks, int32_t rank, const uint8_t* id128) {
    if (!ctx || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return AB200_E_ARGUMENT;
    if (ctx->comm) return AB200_E_ARGUMENT;
    NcclApi* nc = nccl_api(&ctx->err);
    if (!nc) return AB200_E_CUDA;
    CTX_CK(cudaSetDevice(ctx->device));
    ncclUniqueId id; memcpy(&id, id128, 128);
    NCCL_CK(nc->CommInitRank(&ctx->comm, n_ranks, id, rank));
    ctx->n_ranks = n_ranks; ctx->rank = rank;
    return AB200_OK;
}
extern "C" void ab200_comm_destroy(ab200_ctx* ctx) {
    if (!ctx || !ctx->comm) return;
    if (NcclApi* nc = nccl_api(0)) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); nc->CommDestroy(ctx->comm); }
    ctx->comm = 0; ctx->n_ranks = 1; ctx->rank = 0;
}
extern "C" int ab200_comm_info(const ab200_ctx* ctx, int32_t* n_ranks, int32_t* rank) {
    if (!ctx) return AB200_E_ARGUMENT;
    if (n_ranks) *n_ranks = ctx->n_ranks;
    if (rank) *rank = ctx->rank;
    return AB200_OK;
}
// Union over the ranks of this rank's sorted duplicate-free keys (device array `mine`, `cnt` keys): all-gather of the counts,
// all-gather of the payloads padded to the largest count, radix sort of the concatenation and removal of duplicates, all on
// the device; the union is copied to `dst`. Every rank must call it (it is a collective).
static int exchange_keys(ab200_ctx* ctx, const uint64_t* mine, unsigned long long cnt, int total_bits, std::vector<uint64_t>* dst, ab200_status& stl) {
    NcclApi* nc = nccl_api(&ctx->err);
    if (!nc) return AB200_E_CUDA;
    const int n = ctx->n_ranks;
    CTX_CK(ctx->d_counts.ensure(2 * (size_t)n + 2));
    CTX_CK(cudaMemcpyAsync(ctx->d_counts.p + n, &cnt, sizeof(cnt), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_CK(nc->AllGather(ctx->d_counts.p + n, ctx->d_counts.p, 1, ncclUint64, ctx->comm, ctx->stream));
    std::vector<unsigned long long> counts(n);
    CTX_CK(cudaMemcpyAsync(counts.data(), ctx->d_counts.p, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CTX_CK(cudaStreamSynchronize(ctx->stream));
    unsigned long long mx = 0; for (int i = 0; i < n; ++i) if (counts[i] > mx) mx = counts[i];
    if (mx == 0) { if (dst) dst->clear(); return AB200_OK; }
    const size_t total = (size_t)mx * n;
    CTX_CK(ctx->d_gather.ensure(2 * total + mx));            // [send | gathered | sort scratch shares d_tmp]
    uint64_t* send = ctx->d_gather.p; uint64_t* all = ctx->d_gather.p + mx; uint64_t* out = all + total;
    char eb[256]; eb[0] = 0;
    if (dev_launch_fill(ctx->stream, send, mx, ~0ull, eb, sizeof(eb)) != 0) { ctx->err = eb; return AB200_E_CUDA; }
    if (cnt) CTX_CK(cudaMemcpyAsync(send, mine, cnt * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    NCCL_CK(nc->AllGather(send, all, mx, ncclUint64, ctx->comm, ctx->stream));
    // sort the concatenation (padding = all ones sorts last) and drop duplicates
    CTX_CK(ctx->d_tmp.ensure(total));
    size_t nblocks = (total + dev_sort_tile() - 1) / dev_sort_tile();
    CTX_CK(ctx->d_hist.ensure(256 * nblocks));
    int in_tmp = 0;
    if (dev_radix_sort(ctx->stream, all, ctx->d_tmp.p, total, 64, ~0ull, ctx->d_hist.p, &in_tmp, eb, sizeof(eb)) != 0) { ctx->err = eb; return AB200_E_CUDA; }
    (void)total_bits;
    const uint64_t* sorted = in_tmp ? ctx->d_tmp.p : all;
    CTX_CK(ctx->d_pos.ensure(total + 1));
    if (dev_unique_sorted(ctx->stream, sorted, total, ctx->d_pos.p, out, eb, sizeof(eb)) != 0) { ctx->err = eb; return AB200_E_CUDA; }
    stl.launches += 3 * 8 + 4;
    unsigned last_pos = 0; uint64_t last_key = 0, prev_key = 0;
    CTX_CK(cudaMemcpyAsync(&last_pos, ctx->d_pos.p + (total - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    CTX_CK(cudaMemcpyAsync(&last_key, sorted + (total - 1), sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (total > 1) CTX_CK(cudaMemcpyAsync(&prev_key, sorted + (total - 2), sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CTX_CK(cudaStreamSynchronize(ctx->stream));
    const bool last_kept = last_key != ~0ull && (total == 1 || last_key != prev_key);
    const size_t uniq = (size_t)last_pos + (last_kept ? 1 : 0);
    if (dst) {
        dst->resize(uniq);
        if (uniq) CTX_CK(cudaMemcpy(dst->data(), out, uniq * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    }
    return AB200_OK;
}
// All ranks agree on the outcome of a sharded lap: the smallest non-zero code wins on every rank (first error wins, as in
// include/workers.h:129-139, and no rank is left waiting in a collective).
static int agree_on_status(ab200_ctx* ctx, int code, int* agreed) {
    NcclApi* nc = nccl_api(&ctx->err);
    if (!nc) return AB200_E_CUDA;
    const int n = ctx->n_ranks;
    CTX_CK(ctx->d_counts.ensure(2 * (size_t)n + 2));
    unsigned long long c = (unsigned long long)code;
    CTX_CK(cudaMemcpyAsync(ctx->d_counts.p + n, &c, sizeof(c), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_CK(nc->AllGather(ctx->d_counts.p + n, ctx->d_counts.p, 1, ncclUint64, ctx->comm, ctx->stream));
    std::vector<unsigned long long> codes(n);
    CTX_CK(cudaMemcpyAsync(codes.data(), ctx->d_counts.p, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CTX_CK(cudaStreamSynchronize(ctx->stream));
    *agreed = 0;
    for (int i = 0; i < n; ++i) if (codes[i] != 0) { *agreed = (int)codes[i]; break; }
    return AB200_OK;
}

// ---- the lap ---------------------------------------------------------------------------------------------------------
static int map_status(int st) { return st; }   // device ST_* codes coincide with AB200_E_* 1..9

struct LapOut {
    std::vector<uint64_t>* reach_keys; std::vector<uint64_t>* fin_keys;   // sorted unique (may be null)
    ab200_boxlist* reach_boxes; ab200_boxlist* fin_boxes;                 // (may be null)
    const ab200_enclosures* enc_in;      // (may be null) the evolution of unit i starts from enclosure i instead of box i
    ab200_enclosures* enc_out;           // (may be null) receives the final sets as enclosures, in unit order
    int enc_cap;                         // words of enclosure output per unit
};

static int run_batches(ab200_ctx* ctx, uint64_t n_cells, const int32_t* loc, const double* boxes, double T, int32_t max_events,
                       int ignore_activations, int direction, LapOut& out, ab200_status* status) {
    if (!ctx || !ctx->have_system) return AB200_E_ARGUMENT;
    const bool want_keys = out.reach_keys || out.fin_keys;
    if (want_keys && !ctx->have_layout) { ctx->err = "ab200_set_accuracy must be called first"; return AB200_E_ARGUMENT; }
    { int rc_ = upload_system(ctx, false); if (rc_ != AB200_OK) return rc_; }
    const ab200_settings& S = ctx->settings;
    const int stride = AB200_BOX_STRIDE;
    ab200_status stl; memset(&stl, 0, sizeof(stl)); stl.cell = -1;
    for (uint64_t i = 0; i < n_cells; ++i) if (loc[i] < 0 || loc[i] >= ctx->sys.desc.n_locs) { ctx->err = "bad location index"; return AB200_E_ARGUMENT; }
    EvolveParams P;
    P.T = T; P.max_events = max_events; P.ignore_activations = ignore_activations ? 1 : 0; P.max_boxes = S.max_boxes;
    P.temporal_order = S.temporal_order; P.ratio = S.maximum_enclosure_widths_ratio; P.spacial_sweep = S.sweep_threshold;
    P.spacial_order = S.spacial_order; P.max_working_sets = S.max_working_sets > 0 ? S.max_working_sets : 64; P.lower = S.semantics ? 1 : 0; P.backward = direction ? 1 : 0;
    const size_t table_n = want_keys ? ((size_t)1 << S.hash_log2) : 0;
    std::vector<uint64_t> acc_keys[2];
    if (out.reach_boxes) { out.reach_boxes->offsets.assign(1, 0); }
    if (out.fin_boxes) { out.fin_boxes->offsets.assign(1, 0); }
    if (want_keys) {
        CTX_CK(ctx->d_table.ensure(table_n * 2));
        CTX_CK(ctx->d_counters.ensure(12));
        DEV_CK(dev_launch_fill(ctx->stream, ctx->d_table.p, table_n * 2, ~0ull, eb_, sizeof(eb_))); stl.launches += 1;
        CTX_CK(cudaMemsetAsync(ctx->d_counters.p, 0, 8 * sizeof(unsigned long long), ctx->stream));
    }
    const bool wide_sys = ctx->sys.desc.loc[0].n > 2;
    if (out.enc_in) {
        // enclosure hand-over: all records once, the per-unit word offsets beside them
        if (out.enc_in->locs.size() != n_cells) { ctx->err = "one enclosure per unit expected"; return AB200_E_ARGUMENT; }
        CTX_CK(ctx->d_enc_in.ensure(out.enc_in->words.size() + 1)); CTX_CK(ctx->d_enc_off.ensure(n_cells + 1));
        CTX_CK(cudaMemcpyAsync(ctx->d_enc_in.p, out.enc_in->words.data(), out.enc_in->words.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        CTX_CK(cudaMemcpyAsync(ctx->d_enc_off.p, out.enc_in->offsets.data(), n_cells * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (out.enc_out) { out.enc_out->offsets.clear(); out.enc_out->words.clear(); out.enc_out->locs.clear(); out.enc_out->unit_offsets.assign(1, 0); out.enc_out->key_words = wide_sys ? 2 : 1; }
    (void)wide_sys;
    const bool hand_over = out.enc_in || out.enc_out;
    float evolve_ms = 0, raster_ms = 0;
    std::chrono::steady_clock::time_point t_start = std::chrono::steady_clock::now();
    for (uint64_t b0 = 0; b0 < n_cells; b0 += (uint64_t)S.batch_cells) {
        if (S.ttl_seconds > 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() > S.ttl_seconds) {
            ctx->err = "time to live exceeded"; if (status) *status = stl; return AB200_E_TIMEOUT; }
        const int nb = (int)std::min<uint64_t>(S.batch_cells, n_cells - b0);
        CTX_CK(ctx->d_loc.ensure(nb)); CTX_CK(ctx->d_box.ensure((size_t)nb * stride));
        CTX_CK(ctx->d_flops.ensure(nb)); CTX_CK(ctx->d_nreach.ensure(nb)); CTX_CK(ctx->d_nfin.ensure(nb)); CTX_CK(ctx->d_status.ensure(nb)); CTX_CK(ctx->d_steps.ensure(nb)); CTX_CK(ctx->d_peak.ensure(nb));
        CTX_CK(ctx->d_reach.ensure((size_t)nb * S.max_boxes * stride)); CTX_CK(ctx->d_fin.ensure((size_t)nb * S.max_boxes * stride));
        CTX_CK(ctx->d_rloc.ensure((size_t)nb * S.max_boxes)); CTX_CK(ctx->d_floc.ensure((size_t)nb * S.max_boxes));
        // Build selection: set models have up to max(2n+1,3n) parameters -- one-word keys hold 7, two-word keys 15.
        // AB200_TUNE=kw2 forces the two-word build for a system that would fit one word (test hook).
        const bool wide = ctx->sys.desc.loc[0].n > 2;
        const EvolveBuild* build = find_build(wide ? "kw2" : "kw1");
        // AB200_TUNE=lane / warp overrides the size rule for one-word systems (test hook).
        bool use_lane = !wide && nb >= AB_LANE_THRESHOLD && P.max_events < 64 && !hand_over;      // (enclosure records are a warp-build feature)
        if (const char* tune = getenv("AB200_TUNE")) {
            if (!strcmp(tune, "lane") && !wide) use_lane = !hand_over;
            else if (!strcmp(tune, "warp")) use_lane = false;
            else {
                if (!find_build(tune) || (wide && strcmp(tune, "kw2"))) { ctx->err = std::string("AB200_TUNE: no such build for this system: ") + tune; return AB200_E_ARGUMENT; }
                build = find_build(tune); use_lane = false;
            }
        }
        {   // enclosure records carry the MultiIndex keys of the build that wrote them
            const int kw = strcmp(build->name, "kw2") ? 1 : 2;
            if (out.enc_in && out.enc_in->key_words != kw) { ctx->err = "enclosures were written by a build with another key width"; return AB200_E_ARGUMENT; }
            if (out.enc_out) out.enc_out->key_words = kw;
        }
        CTX_CK(cudaMemcpyAsync(ctx->d_loc.p, loc + b0, (size_t)nb * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        CTX_CK(cudaMemcpyAsync(ctx->d_box.p, boxes + b0 * stride, (size_t)nb * stride * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        std::vector<int32_t> h_status(nb), h_steps(nb), h_peak(nb), h_nreach(nb), h_nfin(nb);
        float ms = 0;
        for (int pass = 0; pass < 2; ++pass) {
            CTX_CK(cudaEventRecord(ctx->ev0, ctx->stream));
            if (use_lane) {
                const int lane_words = S.arena_words < AB_LANE_ARENA_WORDS ? S.arena_words : AB_LANE_ARENA_WORDS;
                const int nwarps = (nb + 31) / 32;
                CTX_CK(ctx->d_arena.ensure((size_t)nwarps * 32 * lane_words));
                EvolveParams PL = P; if (PL.max_working_sets > 12) PL.max_working_sets = 12;
                DEV_CK(ab200::lane::dev_launch_evolve_v(ctx->stream, nb, ctx->d_loc.p, ctx->d_box.p, &PL, ctx->d_arena.p, lane_words,
                                                       ctx->d_reach.p, ctx->d_rloc.p, ctx->d_nreach.p, ctx->d_fin.p, ctx->d_floc.p, ctx->d_nfin.p,
                                                       ctx->d_status.p, ctx->d_steps.p, ctx->d_peak.p, eb_, sizeof(eb_)));
            } else {
                int slots; DEV_CK((slots = build->capacity(eb_, sizeof(eb_))) < 0);
                if (const char* e = getenv("AB200_SLOTS_PER_SM")) {     // development hook: fewer resident cells per SM than the occupancy allows
                    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
                    if (atoi(e) > 0 && sms * atoi(e) < slots) slots = sms * atoi(e);
                }
                if (slots > nb + 3) slots = nb + 3;                     // (a CTA may hold up to four cell slots)
                CTX_CK(ctx->d_arena.ensure((size_t)(slots + 4) * S.arena_words));
                CTX_CK(cudaMemsetAsync(ctx->d_next.p, 0, sizeof(int), ctx->stream));
                if (out.enc_out) { CTX_CK(ctx->d_enc_out.ensure((size_t)nb * (size_t)out.enc_cap)); CTX_CK(ctx->d_enc_used.ensure(nb)); }
                DEV_CK(build->launch(ctx->stream, nb, slots, ctx->d_loc.p, ctx->d_box.p, &P, ctx->d_arena.p, S.arena_words, ctx->d_binom.p, ctx->d_next.p,
                                     ctx->d_reach.p, ctx->d_rloc.p, ctx->d_nreach.p, ctx->d_fin.p, ctx->d_floc.p, ctx->d_nfin.p,
                                     ctx->d_status.p, ctx->d_steps.p, ctx->d_peak.p, ctx->d_flops.p,
                                     out.enc_in ? ctx->d_enc_in.p : 0, out.enc_in ? ctx->d_enc_off.p + b0 : 0,
                                     out.enc_out ? ctx->d_enc_out.p : 0, out.enc_cap, out.enc_out ? ctx->d_enc_used.p : 0, eb_, sizeof(eb_)));
            }
            stl.launches += 1;
            CTX_CK(cudaEventRecord(ctx->ev1, ctx->stream));
            CTX_CK(cudaMemcpyAsync(h_status.data(), ctx->d_status.p, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            CTX_CK(cudaStreamSynchronize(ctx->stream));
            float ms1 = 0; cudaEventElapsedTime(&ms1, ctx->ev0, ctx->ev1); ms += ms1;
            // the lane build has fixed device limits (scratch words and pending working sets per cell) and no exp / log / sin / cos:
            // a batch that hits one is run again by the warp build
            // warp builds: bits 8.. of the status word carry the cell's count of libm-sensitive series calls (ab200_series.cuh)
            if (!use_lane) for (int i = 0; i < nb; ++i) { stl.libm_sensitive += (h_status[i] >> 8) & 0xFFFFF; h_status[i] &= 0xFF; }
            bool retry = false;
            if (use_lane) for (int i = 0; i < nb; ++i) if (h_status[i] == ST_ARENA || h_status[i] == ST_WORKING_SETS || h_status[i] == ST_UNSUPPORTED) { retry = true; break; }
            stl.build = use_lane ? 2 : (wide ? 1 : 0);
            if (!retry) break;
            use_lane = false;
        }
        const char* dump_steps = getenv("AB200_DUMP_STEPS");      // profiling aid: per-cell _evolution_step counts of the last lap
        CTX_CK(cudaMemcpyAsync(h_steps.data(), ctx->d_steps.p, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CTX_CK(cudaMemcpyAsync(h_peak.data(), ctx->d_peak.p, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CTX_CK(cudaMemcpyAsync(h_nreach.data(), ctx->d_nreach.p, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CTX_CK(cudaMemcpyAsync(h_nfin.data(), ctx->d_nfin.p, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        std::vector<double> h_flops(stl.build == 2 ? 0 : nb);
        if (!h_flops.empty()) CTX_CK(cudaMemcpyAsync(h_flops.data(), ctx->d_flops.p, nb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CTX_CK(cudaStreamSynchronize(ctx->stream));
        for (size_t i = 0; i < h_flops.size(); ++i) stl.alg_flops += h_flops[i];
        evolve_ms += ms;
        if (dump_steps) { if (FILE* f = fopen(dump_steps, b0 == 0 ? "wb" : "ab")) { fwrite(h_steps.data(), sizeof(int32_t), (size_t)nb, f); fclose(f); } }
        for (int i = 0; i < nb; ++i) {
            stl.steps += h_steps[i];
            if (h_peak[i] > stl.arena_peak) stl.arena_peak = h_peak[i];
            if (h_status[i] != ST_OK && stl.code == AB200_OK) { stl.code = map_status(h_status[i]); stl.cell = (int64_t)(b0 + i); }
        }
        if (stl.code != AB200_OK) {
            // first failure aborts the lap and discards its results (include/workers.h:129-139,159-166)
            char msg[160]; snprintf(msg, sizeof(msg), "cell %lld failed with code %d", (long long)stl.cell, stl.code);
            ctx->err = msg; stl.evolve_ms = evolve_ms; if (status) *status = stl; return stl.code;
        }
        if (out.enc_out) {
            // the final sets of every unit as enclosure records, in unit order and, within a unit, in the order they were emitted
            std::vector<int32_t> used(nb);
            CTX_CK(cudaMemcpy(used.data(), ctx->d_enc_used.p, nb * sizeof(int32_t), cudaMemcpyDeviceToHost));
            std::vector<uint64_t> rec;
            for (int i = 0; i < nb; ++i) {
                rec.resize((size_t)used[i]);
                if (used[i] > 0) CTX_CK(cudaMemcpy(rec.data(), ctx->d_enc_out.p + (size_t)i * (size_t)out.enc_cap, (size_t)used[i] * sizeof(uint64_t), cudaMemcpyDeviceToHost));
                for (size_t o = 0; o < rec.size(); ) {
                    const size_t len = (size_t)(rec[o + 1] & 0xFFFFFFFFu);
                    if (len < 2 || o + len > rec.size()) { ctx->err = "malformed enclosure record"; return AB200_E_CUDA; }
                    out.enc_out->offsets.push_back(out.enc_out->words.size());
                    out.enc_out->locs.push_back((int32_t)(rec[o] >> 32));
                    out.enc_out->words.insert(out.enc_out->words.end(), rec.begin() + o, rec.begin() + o + len);
                    o += len;
                }
                out.enc_out->unit_offsets.push_back(out.enc_out->locs.size());
            }
        }
        if (out.reach_boxes || out.fin_boxes) {
            for (int which = 0; which < 2; ++which) {
                ab200_boxlist* bl = which ? out.fin_boxes : out.reach_boxes;
                if (!bl) continue;
                const std::vector<int32_t>& cnt = which ? h_nfin : h_nreach;
                std::vector<double> hb((size_t)nb * S.max_boxes * stride); std::vector<int32_t> hl((size_t)nb * S.max_boxes);
                CTX_CK(cudaMemcpy(hb.data(), which ? ctx->d_fin.p : ctx->d_reach.p, hb.size() * sizeof(double), cudaMemcpyDeviceToHost));
                CTX_CK(cudaMemcpy(hl.data(), which ? ctx->d_floc.p : ctx->d_rloc.p, hl.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
                for (int i = 0; i < nb; ++i) {
                    for (int k = 0; k < cnt[i]; ++k) {
                        const double* src = &hb[((size_t)i * S.max_boxes + k) * stride];
                        bl->boxes.insert(bl->boxes.end(), src, src + stride);
                        bl->locs.push_back(hl[(size_t)i * S.max_boxes + k]);
                    }
                    bl->offsets.push_back(bl->locs.size());
                }
            }
        }
        if (want_keys) {
            CTX_CK(cudaEventRecord(ctx->ev0, ctx->stream));
            const uint64_t mask = (uint64_t)table_n - 1;
            DEV_CK(dev_launch_raster(ctx->stream, nb, S.max_boxes, ctx->d_reach.p, ctx->d_rloc.p, ctx->d_nreach.p, ctx->K, ctx->sys.desc.n_locs,
                                     ctx->d_table.p, mask, ctx->d_counters.p, eb_, sizeof(eb_)));
            DEV_CK(dev_launch_raster(ctx->stream, nb, S.max_boxes, ctx->d_fin.p, ctx->d_floc.p, ctx->d_nfin.p, ctx->K, ctx->sys.desc.n_locs,
                                     ctx->d_table.p + table_n, mask, ctx->d_counters.p + 3, eb_, sizeof(eb_))); stl.launches += 2;
            CTX_CK(cudaEventRecord(ctx->ev1, ctx->stream));
            CTX_CK(cudaStreamSynchronize(ctx->stream));
            cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); raster_ms += ms;
        }
    }
    if (want_keys) {
        unsigned long long hc[8];
        CTX_CK(cudaMemcpy(hc, ctx->d_counters.p, sizeof(hc), cudaMemcpyDeviceToHost));
        stl.keys_emitted = (int64_t)(hc[2] + hc[5]);
        if (hc[0] || hc[3]) { ctx->err = "cell de-duplication table full: raise ab200_settings.hash_log2"; stl.evolve_ms = evolve_ms; stl.raster_ms = raster_ms; if (status) *status = stl; return AB200_E_KEYSPACE; }
        if (hc[1] || hc[4]) { ctx->err = "a reached cell lies outside the packed-key window"; stl.evolve_ms = evolve_ms; stl.raster_ms = raster_ms; if (status) *status = stl; return AB200_E_KEYSPACE; }
    }
    stl.evolve_ms = evolve_ms; stl.raster_ms = raster_ms;
    if (status) *status = stl;
    return AB200_OK;
}

// Second half of a lap: the de-duplication tables of the two sets -> sorted key arrays (on the host; through the NCCL exchange
// when the lap was sharded over ranks, so that every rank gets the union).
static int lap_keys(ab200_ctx* ctx, LapOut& out, bool sharded, ab200_status* status) {
    const ab200_settings& S = ctx->settings;
    const size_t table_n = (size_t)1 << S.hash_log2;
    ab200_status stl; if (status) stl = *status; else memset(&stl, 0, sizeof(stl));
    float raster_ms = (float)stl.raster_ms;
    {
        CTX_CK(cudaEventRecord(ctx->ev0, ctx->stream));
        int total_bits = ctx->K.loc_bits; for (int i = 0; i < ctx->K.dim; ++i) total_bits += ctx->K.bits[i];
        for (int which = 0; which < 2; ++which) {
            std::vector<uint64_t>* dst = which ? out.fin_keys : out.reach_keys;
            { const unsigned long long init3[3] = { 0ull, 0ull, ~0ull };        // count, OR of the keys, AND of the keys
              CTX_CK(cudaMemcpyAsync(ctx->d_counters.p + 8, init3, sizeof(init3), cudaMemcpyHostToDevice, ctx->stream)); }
            CTX_CK(ctx->d_keys.ensure(table_n)); CTX_CK(ctx->d_tmp.ensure(table_n));
            DEV_CK(dev_launch_compact(ctx->stream, ctx->d_table.p + (size_t)which * table_n, table_n, ctx->d_keys.p, ctx->d_counters.p + 8, eb_, sizeof(eb_))); stl.launches += 1;
            unsigned long long cnt3[3] = { 0ull, 0ull, 0ull };
            CTX_CK(cudaMemcpyAsync(cnt3, ctx->d_counters.p + 8, sizeof(cnt3), cudaMemcpyDeviceToHost, ctx->stream));
            CTX_CK(cudaStreamSynchronize(ctx->stream));
            const unsigned long long cnt = cnt3[0];
            const unsigned long long varying = cnt ? (cnt3[1] ^ cnt3[2]) : 0ull;       // bits in which the keys of this set differ
            int in_tmp = 0;
            if (cnt > 1) {
                size_t nblocks = (cnt + dev_sort_tile() - 1) / dev_sort_tile();
                CTX_CK(ctx->d_hist.ensure(256 * nblocks));
                DEV_CK(dev_radix_sort(ctx->stream, ctx->d_keys.p, ctx->d_tmp.p, cnt, total_bits, varying, ctx->d_hist.p, &in_tmp, eb_, sizeof(eb_)));
                for (int shift = 0; shift < total_bits; shift += 8) if ((varying >> shift) & 0xFFull) stl.launches += 3;
            }
            if (sharded) {
                int rc_ = exchange_keys(ctx, in_tmp ? ctx->d_tmp.p : ctx->d_keys.p, cnt, total_bits, dst, stl);
                if (rc_ != AB200_OK) return rc_;
            } else if (dst) {
                dst->resize(cnt);
                if (cnt) CTX_CK(cudaMemcpyAsync(dst->data(), in_tmp ? ctx->d_tmp.p : ctx->d_keys.p, cnt * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
            }
            CTX_CK(cudaStreamSynchronize(ctx->stream));
        }
        CTX_CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CTX_CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); raster_ms += ms;
    }
    stl.raster_ms = raster_ms;
    if (status) *status = stl;
    return AB200_OK;
}

// One lap. With a communicator (ab200_comm_init) every rank passes the SAME frontier; blocks of 64 cells are dealt round-robin
// to the ranks, each rank evolves and rasterises its share, the ranks agree on the outcome, and the sorted keys
// are exchanged on the device, so every rank returns the sets of the whole lap.
static int run_lap(ab200_ctx* ctx, uint64_t n_cells, const int32_t* loc, const double* boxes, double T, int32_t max_events,
                   int ignore_activations, int direction, LapOut& out, ab200_status* status) {
    ab200_status stl; memset(&stl, 0, sizeof(stl)); stl.cell = -1;
    const bool want_keys = out.reach_keys || out.fin_keys;
    const bool sharded = ctx && ctx->comm && ctx->n_ranks > 1 && want_keys;
    int rc;
    if (!sharded) rc = run_batches(ctx, n_cells, loc, boxes, T, max_events, ignore_activations, direction, out, &stl);
    else {
        const int stride = AB200_BOX_STRIDE;
        std::vector<int32_t> ml; std::vector<double> mb; std::vector<uint64_t> index;
        for (uint64_t i = 0; i < n_cells; ++i) if ((int)((i / 64) % (uint64_t)ctx->n_ranks) == ctx->rank) {
            ml.push_back(loc[i]); mb.insert(mb.end(), boxes + i * stride, boxes + (i + 1) * stride); index.push_back(i);
        }
        rc = run_batches(ctx, ml.size(), ml.data(), mb.data(), T, max_events, ignore_activations, direction, out, &stl);
        if (stl.cell >= 0 && (size_t)stl.cell < index.size()) stl.cell = (int64_t)index[(size_t)stl.cell];
        int agreed = 0;
        int rc2 = agree_on_status(ctx, rc, &agreed);
        if (rc2 != AB200_OK) { if (status) *status = stl; return rc2; }
        if (agreed != 0 && rc == AB200_OK) ctx->err = "another rank failed its share of the lap";
        rc = agreed;
    }
    if (rc == AB200_OK && want_keys) rc = lap_keys(ctx, out, sharded, &stl);
    if (status) *status = stl;
    return rc;
}

extern "C" int ab200_upper_reach_evolve(ab200_ctx* ctx, uint64_t n_cells, const int32_t* loc, const double* boxes,
                                        double t_continuous, int32_t max_events, int ignore_activations, int direction,
                                        ab200_cellset** reach_out, ab200_cellset** evolve_out, ab200_status* status) {
    if (!ctx || !reach_out || !evolve_out || (n_cells && (!loc || !boxes))) return AB200_E_ARGUMENT;
    *reach_out = 0; *evolve_out = 0;
    ab200_cellset* r = new ab200_cellset(); ab200_cellset* e = new ab200_cellset();
    LapOut out; out.reach_keys = &r->keys; out.fin_keys = &e->keys; out.reach_boxes = 0; out.fin_boxes = 0; out.enc_in = 0; out.enc_out = 0; out.enc_cap = 0;
    int rc = run_lap(ctx, n_cells, loc, boxes, t_continuous, max_events, ignore_activations, direction, out, status);
    if (rc != AB200_OK) { delete r; delete e; return rc; }
    *reach_out = r; *evolve_out = e;
    return AB200_OK;
}
extern "C" int ab200_cellset_view(const ab200_cellset* cs, const uint64_t** keys, uint64_t* count) {
    if (!cs) return AB200_E_ARGUMENT;
    if (keys) *keys = cs->keys.empty() ? 0 : cs->keys.data();
    if (count) *count = cs->keys.size();
    return AB200_OK;
}
extern "C" void ab200_cellset_free(ab200_cellset* cs) { delete cs; }

extern "C" int ab200_reach_evolve_boxes(ab200_ctx* ctx, uint64_t n_cells, const int32_t* loc, const double* boxes,
                                        double t_continuous, int32_t max_events, int ignore_activations, int direction,
                                        ab200_boxlist** reach_out, ab200_boxlist** final_out, ab200_status* status) {
    if (!ctx || !reach_out || !final_out || (n_cells && (!loc || !boxes))) return AB200_E_ARGUMENT;
    *reach_out = 0; *final_out = 0;
    ab200_boxlist* r = new ab200_boxlist(); ab200_boxlist* f = new ab200_boxlist();
    LapOut out; out.reach_keys = 0; out.fin_keys = 0; out.reach_boxes = r; out.fin_boxes = f; out.enc_in = 0; out.enc_out = 0; out.enc_cap = 0;
    int rc = run_lap(ctx, n_cells, loc, boxes, t_continuous, max_events, ignore_activations, direction, out, status);
    if (rc != AB200_OK) { delete r; delete f; return rc; }
    *reach_out = r; *final_out = f;
    return AB200_OK;
}
extern "C" int ab200_boxlist_view(const ab200_boxlist* bl, const uint64_t** offsets, const double** boxes, const int32_t** locs) {
    if (!bl) return AB200_E_ARGUMENT;
    if (offsets) *offsets = bl->offsets.data();
    if (boxes) *boxes = bl->boxes.empty() ? 0 : bl->boxes.data();
    if (locs) *locs = bl->locs.empty() ? 0 : bl->locs.data();
    return AB200_OK;
}
extern "C" void ab200_boxlist_free(ab200_boxlist* bl) { delete bl; }

// reach_evolve with enclosure hand-over: EvolverInterface::reach_evolve(enclosure,time,semantics) (include/evolver_interface.h:95-124)
// as LowerReachEpsilonWorker::_compute calls it (include/workers.h:264-303) -- every unit starts from a TaylorSet (a final set of
// an earlier call) or from a box, and its final sets come back as TaylorSets as well as boxes.
extern "C" int ab200_reach_evolve_enclosures(ab200_ctx* ctx, uint64_t n_units, const int32_t* loc, const double* boxes, const ab200_enclosures* initial,
                                             double t_continuous, int32_t max_events, int ignore_activations, int direction, int32_t max_enclosure_words,
                                             ab200_boxlist** reach_out, ab200_boxlist** final_out, ab200_enclosures** final_enclosures, ab200_status* status) {
    if (!ctx || !reach_out || !final_out || !final_enclosures) return AB200_E_ARGUMENT;
    if (!initial && n_units && (!loc || !boxes)) return AB200_E_ARGUMENT;
    *reach_out = 0; *final_out = 0; *final_enclosures = 0;
    std::vector<int32_t> locs; std::vector<double> zero;
    if (initial) {
        // the units are the enclosures; their locations come from the records and the boxes are not read
        n_units = initial->locs.size(); locs = initial->locs; zero.assign((size_t)n_units * AB200_BOX_STRIDE, 0.0);
        loc = locs.data(); boxes = zero.data();
    }
    ab200_boxlist* r = new ab200_boxlist(); ab200_boxlist* f = new ab200_boxlist(); ab200_enclosures* e = new ab200_enclosures();
    LapOut out; out.reach_keys = 0; out.fin_keys = 0; out.reach_boxes = r; out.fin_boxes = f;
    out.enc_in = initial; out.enc_out = e; out.enc_cap = max_enclosure_words > 0 ? max_enclosure_words : 16384;
    int rc = run_lap(ctx, n_units, loc, boxes, t_continuous, max_events, ignore_activations, direction, out, status);
    if (rc != AB200_OK) { delete r; delete f; delete e; return rc; }
    *reach_out = r; *final_out = f; *final_enclosures = e;
    return AB200_OK;
}
extern "C" int ab200_enclosures_view(const ab200_enclosures* e, uint64_t* count, const int32_t** locs, const uint64_t** unit_offsets, uint64_t* n_units) {
    if (!e) return AB200_E_ARGUMENT;
    if (count) *count = e->locs.size();
    if (locs) *locs = e->locs.empty() ? 0 : e->locs.data();
    if (unit_offsets) *unit_offsets = e->unit_offsets.empty() ? 0 : e->unit_offsets.data();
    if (n_units) *n_units = e->unit_offsets.empty() ? 0 : e->unit_offsets.size() - 1;
    return AB200_OK;
}
// Model k (a coordinate of the set) of enclosure i, decoded: exps is [n_terms][nargs] (MultiIndex exponents), coefs [n_terms].
// With null exps / coefs only the sizes and the error are returned.
extern "C" int ab200_enclosures_model(const ab200_enclosures* e, uint64_t i, int32_t k, int32_t* nargs, int32_t* n_terms, double* error, uint8_t* exps, double* coefs) {
    if (!e || i >= e->locs.size() || k < 0) return AB200_E_ARGUMENT;
    const uint64_t* q = e->words.data() + e->offsets[i];
    const int n = (int)(q[1] >> 32), tw = e->key_words + 1;
    if (k >= n) return AB200_E_ARGUMENT;
    size_t o = 2;
    for (int m = 0; m < k; ++m) o += 2 + (size_t)(q[o] >> 32) * tw;
    const int nt = (int)(q[o] >> 32), as = (int)(q[o] & 0xFFFFFFFFu);
    if (nargs) *nargs = as;
    if (n_terms) *n_terms = nt;
    if (error) memcpy(error, &q[o + 1], sizeof(double));
    for (int t = 0; t < nt; ++t) {
        const uint64_t* w = q + o + 2 + (size_t)t * tw;
        if (exps) for (int j = 0; j < as; ++j) exps[(size_t)t * as + j] = (uint8_t)(j < 7 ? (w[0] >> (8 * (6 - j))) & 0xFF : (e->key_words > 1 ? (w[1] >> (8 * (14 - j))) & 0xFF : 0));
        if (coefs) memcpy(&coefs[t], &w[e->key_words], sizeof(double));
    }
    return AB200_OK;
}
// The enclosures idx[0..n) of src as a new list (the hand-over keeps a subset: ReachabilityRestriction::filter, _filter_enclosures).
extern "C" int ab200_enclosures_select(const ab200_enclosures* src, uint64_t n, const uint64_t* idx, ab200_enclosures** out) {
    if (!src || !out || (n && !idx)) return AB200_E_ARGUMENT;
    ab200_enclosures* e = new ab200_enclosures(); e->key_words = src->key_words; e->unit_offsets.assign(1, 0);
    for (uint64_t k = 0; k < n; ++k) {
        if (idx[k] >= src->locs.size()) { delete e; return AB200_E_ARGUMENT; }
        const uint64_t* q = src->words.data() + src->offsets[idx[k]];
        e->offsets.push_back(e->words.size()); e->locs.push_back(src->locs[idx[k]]);
        e->words.insert(e->words.end(), q, q + (size_t)(q[1] & 0xFFFFFFFFu));
        e->unit_offsets.push_back(e->locs.size());
    }
    *out = e;
    return AB200_OK;
}
// Enclosures from the caller's own TaylorModels (the reverse of ab200_enclosures_model): what an adapter does with the
// TaylorSets the reference hands to LowerReachEpsilonWorker (include/workers.h:171-200). Terms must come in the order the
// reference's Expansion keeps them (graded, include/multi_index.h:346-358), which is the order the device merges rely on.
extern "C" int ab200_enclosures_create(ab200_ctx* ctx, uint64_t count, const int32_t* locs, const int32_t* nargs, const int32_t* n_terms,
                                       const double* errors, const uint8_t* exps, const double* coefs, ab200_enclosures** out) {
    if (!ctx || !ctx->have_system || !out || (count && (!locs || !nargs || !n_terms || !errors))) return AB200_E_ARGUMENT;
    *out = 0;
    const int kw = ctx->sys.desc.loc[0].n > 2 ? 2 : 1, maxargs = kw == 1 ? 7 : 15;
    ab200_enclosures* e = new ab200_enclosures(); e->key_words = kw; e->unit_offsets.assign(1, 0);
    size_t m = 0, t = 0, x = 0;                  // running model, term and exponent indices
    const uint64_t M = 0x00FFFFFFFFFFFFFFull;
    for (uint64_t i = 0; i < count; ++i) {
        const int l = locs[i], as = nargs[i];
        if (l < 0 || l >= ctx->sys.desc.n_locs || as < 0 || as > maxargs) { delete e; ctx->err = "enclosure: bad location or too many parameters"; return AB200_E_ARGUMENT; }
        const int n = ctx->sys.desc.loc[l].n;
        const size_t start = e->words.size();
        e->offsets.push_back(start); e->locs.push_back(l);
        e->words.push_back(((uint64_t)(uint32_t)l << 32) | (uint32_t)as); e->words.push_back(0);
        for (int k = 0; k < n; ++k, ++m) {
            const int nt = n_terms[m];
            if (nt < 0 || !(errors[m] >= 0.0) || (nt && (!exps || !coefs))) { delete e; ctx->err = "enclosure: bad model"; return AB200_E_ARGUMENT; }
            uint64_t eb; memcpy(&eb, &errors[m], 8);
            e->words.push_back(((uint64_t)(uint32_t)nt << 32) | (uint32_t)as); e->words.push_back(eb);
            uint64_t p0 = 0, p1 = 0;
            for (int q = 0; q < nt; ++q, ++t) {
                uint64_t w0 = 0, w1 = 0; unsigned deg = 0;
                for (int j = 0; j < as; ++j, ++x) {
                    const unsigned ex = exps[x];
                    if (ex > 127) { delete e; ctx->err = "enclosure: exponent above 127"; return AB200_E_ARGUMENT; }
                    deg += ex;
                    if (j < 7) w0 |= (uint64_t)ex << (8 * (6 - j)); else w1 |= (uint64_t)ex << (8 * (14 - j));
                }
                if (deg > 255) { delete e; ctx->err = "enclosure: degree above 255"; return AB200_E_ARGUMENT; }
                w0 |= (uint64_t)deg << 56;
                if (q > 0) {        // strictly ascending in the order of key_less (ab200_tm.cuh)
                    const uint64_t sa = p0 ^ M, sb = w0 ^ M;
                    if (!(sa < sb || (sa == sb && p1 > w1))) { delete e; ctx->err = "enclosure: terms out of order"; return AB200_E_ARGUMENT; }
                }
                p0 = w0; p1 = w1;
                uint64_t cb; memcpy(&cb, &coefs[t], 8);
                e->words.push_back(w0); if (kw > 1) e->words.push_back(w1); e->words.push_back(cb);
            }
        }
        e->words[start + 1] = ((uint64_t)(uint32_t)n << 32) | (uint32_t)(e->words.size() - start);
        e->unit_offsets.push_back(e->locs.size());
    }
    *out = e;
    return AB200_OK;
}
extern "C" void ab200_enclosures_free(ab200_enclosures* e) { delete e; }

extern "C" int ab200_flow_bounds(ab200_ctx* ctx, uint64_t n, const int32_t* loc, const double* boxes, double hmax,
                                 double* h_out, double* bounds_out, int32_t* status_out) {
    if (!ctx || !ctx->have_system || !loc || !boxes || !h_out || !bounds_out) return AB200_E_ARGUMENT;
    { int rc_ = upload_system(ctx, false); if (rc_ != AB200_OK) return rc_; }
    const int stride = AB200_BOX_STRIDE;
    DevBuf<double> d_h, d_b;
    CTX_CK(ctx->d_loc.ensure(n)); CTX_CK(ctx->d_box.ensure(n * stride)); CTX_CK(ctx->d_status.ensure(n));
    cudaError_t e1 = d_h.ensure(n), e2 = d_b.ensure(n * stride);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { d_h.release(); d_b.release(); ctx->err = "cudaMalloc failed"; return AB200_E_CUDA; }
    cudaMemcpyAsync(ctx->d_loc.p, loc, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(ctx->d_box.p, boxes, n * stride * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    char eb[256]; eb[0] = 0;
    int rc = dev_launch_flow_bounds(ctx->stream, (int)n, ctx->d_loc.p, ctx->d_box.p, hmax, d_h.p, d_b.p, ctx->d_status.p, eb, sizeof(eb));
    if (rc == 0) {
        cudaMemcpyAsync(h_out, d_h.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(bounds_out, d_b.p, n * stride * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (status_out) cudaMemcpyAsync(status_out, ctx->d_status.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
    }
    cudaError_t es = cudaStreamSynchronize(ctx->stream);
    d_h.release(); d_b.release();
    if (rc != 0) { ctx->err = eb; return AB200_E_CUDA; }
    if (es != cudaSuccess) { ctx->err = cudaGetErrorString(es); return AB200_E_CUDA; }
    return AB200_OK;
}

// == IntegratorBase::flow_bounds (src/integrator.cpp:44-113)
extern "C" int ab200_integrator_flow_bounds(ab200_ctx* ctx, uint64_t n, const int32_t* loc, const double* boxes, double hmax,
                                            double* h_out, double* bounds_out, int32_t* status_out) {
    if (!ctx || !ctx->have_system || !loc || !boxes || !h_out || !bounds_out) return AB200_E_ARGUMENT;
    { int rc_ = upload_system(ctx, false); if (rc_ != AB200_OK) return rc_; }
    const int stride = AB200_BOX_STRIDE;
    DevBuf<double> d_h, d_b;
    CTX_CK(ctx->d_loc.ensure(n)); CTX_CK(ctx->d_box.ensure(n * stride)); CTX_CK(ctx->d_status.ensure(n));
    cudaError_t e1 = d_h.ensure(n), e2 = d_b.ensure(n * stride);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { d_h.release(); d_b.release(); ctx->err = "cudaMalloc failed"; return AB200_E_CUDA; }
    cudaMemcpyAsync(ctx->d_loc.p, loc, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(ctx->d_box.p, boxes, n * stride * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    char eb[256]; eb[0] = 0;
    int rc = dev_launch_integrator_flow_bounds(ctx->stream, (int)n, ctx->d_loc.p, ctx->d_box.p, hmax, d_h.p, d_b.p, ctx->d_status.p, eb, sizeof(eb));
    if (rc == 0) {
        cudaMemcpyAsync(h_out, d_h.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(bounds_out, d_b.p, n * stride * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (status_out) cudaMemcpyAsync(status_out, ctx->d_status.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
    }
    cudaError_t es = cudaStreamSynchronize(ctx->stream);
    d_h.release(); d_b.release();
    if (rc != 0) { ctx->err = eb; return AB200_E_CUDA; }
    if (es != cudaSuccess) { ctx->err = cudaGetErrorString(es); return AB200_E_CUDA; }
    return AB200_OK;
}

struct ab200_taylor_models { int nargs; std::vector<int32_t> n_terms; std::vector<double> error; std::vector<std::vector<uint8_t> > exps; std::vector<std::vector<double> > coefs; };
// == TaylorIntegrator::flow (src/integrator.cpp:116-152)
extern "C" int ab200_integrator_flow(ab200_ctx* ctx, int32_t loc, const double* box, double hmax, int32_t temporal_order,
                                     double* h_out, double* domain_out, ab200_taylor_models** out) {
    if (!ctx || !ctx->have_system || !box || !out || loc < 0 || loc >= ctx->sys.desc.n_locs || temporal_order < 0) return AB200_E_ARGUMENT;
    *out = 0;
    { int rc_ = upload_system(ctx, false); if (rc_ != AB200_OK) return rc_; }
    const int n = ctx->sys.desc.loc[loc].n;
    if (n + 1 > 7) { ctx->err = "integrator flow supports up to 6 state variables"; return AB200_E_UNSUPPORTED; }
    const int out_words = 1 << 20;
    DevBuf<uint64_t> d_out; DevBuf<double> d_bx;
    CTX_CK(ctx->d_arena.ensure((size_t)ctx->settings.arena_words));
    if (d_out.ensure(out_words) != cudaSuccess || d_bx.ensure(AB200_BOX_STRIDE) != cudaSuccess) { d_out.release(); d_bx.release(); ctx->err = "cudaMalloc failed"; return AB200_E_CUDA; }
    cudaMemcpyAsync(d_bx.p, box, 2 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    char eb[256]; eb[0] = 0;
    int rc = dev_launch_integrator_flow(ctx->stream, loc, d_bx.p, hmax, temporal_order, ctx->d_arena.p, ctx->settings.arena_words, ctx->d_binom.p, d_out.p, out_words, eb, sizeof(eb));
    std::vector<uint64_t> h(3);
    if (rc == 0) cudaMemcpyAsync(h.data(), d_out.p, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t es = cudaStreamSynchronize(ctx->stream);
    if (rc != 0 || es != cudaSuccess) { d_out.release(); d_bx.release(); ctx->err = rc ? eb : cudaGetErrorString(es); return AB200_E_CUDA; }
    double hh; memcpy(&hh, &h[0], 8);
    if (h_out) *h_out = hh;
    if (domain_out) { for (int i = 0; i < n; ++i) { domain_out[2 * i] = box[2 * i]; domain_out[2 * i + 1] = box[2 * i + 1]; } domain_out[2 * n] = -hh; domain_out[2 * n + 1] = hh; }
    if ((int)h[1] != ST_OK) { d_out.release(); d_bx.release(); ctx->err = "integrator flow failed on the device"; return map_status((int)h[1]); }
    std::vector<uint64_t> w(h[2]);
    cudaMemcpy(w.data(), d_out.p, h[2] * sizeof(uint64_t), cudaMemcpyDeviceToHost);
    d_out.release(); d_bx.release();
    ab200_taylor_models* m = new ab200_taylor_models(); m->nargs = n + 1;
    size_t p = 3;
    for (int i = 0; i < n; ++i) {
        const int nt = (int)w[p++]; double e; memcpy(&e, &w[p++], 8);
        m->n_terms.push_back(nt); m->error.push_back(e);
        std::vector<uint8_t> ex((size_t)nt * (n + 1)); std::vector<double> cf(nt);
        for (int t = 0; t < nt; ++t) {
            const uint64_t key = w[p++]; memcpy(&cf[t], &w[p++], 8);
            for (int j = 0; j <= n; ++j) ex[(size_t)t * (n + 1) + j] = (uint8_t)((key >> (8 * (6 - j))) & 0xFF);
        }
        m->exps.push_back(ex); m->coefs.push_back(cf);
    }
    *out = m;
    return AB200_OK;
}
// ---- primitive-level entry points: one TaylorModel operator / function, a batch of Interval operations ---------------
// (include/taylor_model.h operators and functions, include/numeric.h Interval arithmetic; used by the known-answer tests,
// which replay the vectors of the reference's test/test_taylor_model.cpp and test/test_interval.cpp on the device)
extern "C" int ab200_taylor_model_op(ab200_ctx* ctx, int32_t op, int32_t nargs,
                                     int32_t x_terms, const uint8_t* x_exps, const double* x_coefs, double x_error,
                                     int32_t y_terms, const uint8_t* y_exps, const double* y_coefs, double y_error,
                                     int32_t k, int32_t flag, double s, const double* interval, double* range_out, ab200_taylor_models** out) {
    if (!ctx || !out || nargs < 1 || nargs > 7 || x_terms < 0 || y_terms < 0 || (x_terms && (!x_exps || !x_coefs)) || (y_terms && (!y_exps || !y_coefs))) return AB200_E_ARGUMENT;
    *out = 0;
    std::vector<uint64_t> in;
    double iv[2] = { interval ? interval[0] : 0.0, interval ? interval[1] : 0.0 };
    uint64_t u; 
    in.push_back((uint64_t)op); in.push_back((uint64_t)nargs); in.push_back((uint64_t)(int64_t)k); in.push_back((uint64_t)(int64_t)flag);
    memcpy(&u, &s, 8); in.push_back(u); memcpy(&u, &iv[0], 8); in.push_back(u); memcpy(&u, &iv[1], 8); in.push_back(u);
    const int32_t nt[2] = { x_terms, y_terms }; const uint8_t* ex[2] = { x_exps, y_exps }; const double* cf[2] = { x_coefs, y_coefs }; const double er[2] = { x_error, y_error };
    for (int q = 0; q < 2; ++q) {
        in.push_back((uint64_t)nt[q]); memcpy(&u, &er[q], 8); in.push_back(u);
        for (int t = 0; t < nt[q]; ++t) {
            uint64_t key = 0; unsigned deg = 0;
            for (int j = 0; j < nargs; ++j) { const unsigned e = ex[q][(size_t)t * nargs + j]; if (e > 127) return AB200_E_ARGUMENT; deg += e; key |= (uint64_t)e << (8 * (6 - j)); }
            if (deg > 255) return AB200_E_ARGUMENT;
            key |= (uint64_t)deg << 56;
            in.push_back(key); memcpy(&u, &cf[q][t], 8); in.push_back(u);
        }
    }
    const int out_words = 1 << 18;
    DevBuf<uint64_t> d_in, d_out;
    CTX_CK(ctx->d_arena.ensure((size_t)ctx->settings.arena_words));
    if (d_in.ensure(in.size()) != cudaSuccess || d_out.ensure(out_words) != cudaSuccess) { d_in.release(); d_out.release(); ctx->err = "cudaMalloc failed"; return AB200_E_CUDA; }
    cudaMemcpyAsync(d_in.p, in.data(), in.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
    char eb[256]; eb[0] = 0;
    int rc = dev_launch_tm_op(ctx->stream, d_in.p, ctx->d_arena.p, ctx->settings.arena_words, ctx->d_binom.p, d_out.p, out_words, eb, sizeof(eb));
    std::vector<uint64_t> h(4);
    if (rc == 0) cudaMemcpyAsync(h.data(), d_out.p, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t es = cudaStreamSynchronize(ctx->stream);
    if (rc != 0 || es != cudaSuccess) { d_in.release(); d_out.release(); ctx->err = rc ? eb : cudaGetErrorString(es); return AB200_E_CUDA; }
    if ((int)h[0] != ST_OK) { d_in.release(); d_out.release(); ctx->err = "TaylorModel operation failed on the device"; return map_status((int)h[0]); }
    if (range_out) { memcpy(&range_out[0], &h[2], 8); memcpy(&range_out[1], &h[3], 8); }
    std::vector<uint64_t> w(h[1]);
    cudaMemcpy(w.data(), d_out.p, h[1] * sizeof(uint64_t), cudaMemcpyDeviceToHost);
    d_in.release(); d_out.release();
    ab200_taylor_models* m = new ab200_taylor_models(); m->nargs = nargs;
    size_t p = 4;
    const int n = (int)w[p++]; double e; memcpy(&e, &w[p++], 8);
    m->n_terms.push_back(n); m->error.push_back(e);
    std::vector<uint8_t> exo((size_t)n * nargs); std::vector<double> cfo(n);
    for (int t = 0; t < n; ++t) {
        const uint64_t key = w[p++]; memcpy(&cfo[t], &w[p++], 8);
        for (int j = 0; j < nargs; ++j) exo[(size_t)t * nargs + j] = (uint8_t)((key >> (8 * (6 - j))) & 0xFF);
    }
    m->exps.push_back(exo); m->coefs.push_back(cfo);
    *out = m;
    return AB200_OK;
}
// ops[i] = {op, m, a.lower, a.upper, b.lower, b.upper}; results[i] = {lower, upper, status}
extern "C" int ab200_interval_op(ab200_ctx* ctx, int32_t n, const double* ops, double* results) {
    if (!ctx || n < 0 || (n && (!ops || !results))) return AB200_E_ARGUMENT;
    if (n == 0) return AB200_OK;
    DevBuf<double> d_in, d_out;
    if (d_in.ensure((size_t)n * 6) != cudaSuccess || d_out.ensure((size_t)n * 3) != cudaSuccess) { d_in.release(); d_out.release(); ctx->err = "cudaMalloc failed"; return AB200_E_CUDA; }
    cudaMemcpyAsync(d_in.p, ops, (size_t)n * 6 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    char eb[256]; eb[0] = 0;
    int rc = dev_launch_interval_op(ctx->stream, n, d_in.p, d_out.p, eb, sizeof(eb));
    if (rc == 0) cudaMemcpyAsync(results, d_out.p, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t es = cudaStreamSynchronize(ctx->stream);
    d_in.release(); d_out.release();
    if (rc != 0 || es != cudaSuccess) { ctx->err = rc ? eb : cudaGetErrorString(es); return AB200_E_CUDA; }
    return AB200_OK;
}
extern "C" int ab200_taylor_models_view(const ab200_taylor_models* m, int32_t i, int32_t* nargs, int32_t* n_terms, double* error,
                                        const uint8_t** exponents, const double** coefficients) {
    if (!m || i < 0 || i >= (int)m->n_terms.size()) return AB200_E_ARGUMENT;
    if (nargs) *nargs = m->nargs;
    if (n_terms) *n_terms = m->n_terms[i];
    if (error) *error = m->error[i];
    if (exponents) *exponents = m->exps[i].empty() ? 0 : m->exps[i].data();
    if (coefficients) *coefficients = m->coefs[i].empty() ? 0 : m->coefs[i].data();
    return AB200_OK;
}
extern "C" void ab200_taylor_models_free(ab200_taylor_models* m) { delete m; }

static int probe_fp64(ab200_ctx* ctx, int fma, double* gflops) {
    if (!ctx || !gflops) return AB200_E_ARGUMENT;
    CTX_CK(cudaSetDevice(ctx->device));
    const int blocks = 148 * 8, threads = 256, iters = 20000;
    DevBuf<double> scratch;
    if (scratch.ensure((size_t)blocks * threads) != cudaSuccess) { ctx->err = "cudaMalloc failed"; return AB200_E_CUDA; }
    char eb[256];
    dev_fp64_peak(ctx->stream, scratch.p, blocks, threads, 100, fma, eb, sizeof(eb));   // warm-up
    cudaStreamSynchronize(ctx->stream);
    cudaEventRecord(ctx->ev0, ctx->stream);
    dev_fp64_peak(ctx->stream, scratch.p, blocks, threads, iters, fma, eb, sizeof(eb));
    cudaEventRecord(ctx->ev1, ctx->stream);
    cudaError_t es = cudaStreamSynchronize(ctx->stream);
    scratch.release();
    if (es != cudaSuccess) { ctx->err = cudaGetErrorString(es); return AB200_E_CUDA; }
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    double ops = (double)blocks * threads * (double)iters * 16.0;   // 8 accumulators x (1 mul + 1 add), or 8 FMAs = 16 operations
    *gflops = ops / (ms * 1e-3) / 1e9;
    return AB200_OK;
}
extern "C" int ab200_probe_fp64(ab200_ctx* ctx, double* gflops) { return probe_fp64(ctx, 0, gflops); }
extern "C" int ab200_probe_fp64_fma(ab200_ctx* ctx, double* gflops) { return probe_fp64(ctx, 1, gflops); }

// the analyser entry points live in ab200_analyser.cu and use these two accessors
const ab200::HostSystem* ab200_ctx_system(const ab200_ctx* ctx) { return ctx && ctx->have_system ? &ctx->sys : 0; }
// ForwardDiscreteJumpSetChecker::check for a batch of cell boxes on the device (ab200_analyser.cu: forward_jump_set)
int ab200_ctx_jump_check(ab200_ctx* ctx, int loc, int event, size_t n, const double* boxes, int* tri, int* lie_missing) {
    if (!ctx || !ctx->have_system || loc < 0 || loc >= ctx->sys.desc.n_locs || event < 0 || event >= ctx->sys.desc.loc[loc].n_events) return AB200_E_ARGUMENT;
    if (n == 0) return AB200_OK;
    { int rc_ = upload_system(ctx, false); if (rc_ != AB200_OK) return rc_; }
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    int blocks = sms * 8; if ((size_t)blocks > n) blocks = (int)n;
    const int aw = ctx->settings.arena_words;
    CTX_CK(ctx->d_arena.ensure((size_t)blocks * aw));
    CTX_CK(ctx->d_jbox.ensure(n * AB200_BOX_STRIDE)); CTX_CK(ctx->d_jout.ensure(2 * n + 1));
    int32_t* const d_out = ctx->d_jout.p;
    cudaMemcpyAsync(ctx->d_jbox.p, boxes, n * AB200_BOX_STRIDE * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemsetAsync(ctx->d_next.p, 0, sizeof(int), ctx->stream);
    cudaMemsetAsync(d_out + 2 * n, 0, sizeof(int32_t), ctx->stream);
    char eb[256]; eb[0] = 0;
    int rc = dev_launch_jump_check(ctx->stream, (int)n, blocks, loc, event, ctx->d_jbox.p, ctx->d_arena.p, aw, ctx->d_binom.p, ctx->d_next.p,
                                   d_out, d_out + n, d_out + 2 * n, eb, sizeof(eb));
    std::vector<int32_t> h(2 * n + 1);
    if (rc == 0) cudaMemcpyAsync(h.data(), d_out, (2 * n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t es = cudaStreamSynchronize(ctx->stream);
    if (rc != 0 || es != cudaSuccess) { ctx->err = rc ? eb : cudaGetErrorString(es); return AB200_E_CUDA; }
    if (lie_missing && h[2 * n]) *lie_missing = 1;
    for (size_t i = 0; i < n; ++i) {
        tri[i] = h[i];
        if (h[n + i] != ST_OK && !(h[2 * n] && h[n + i] == ST_UNSUPPORTED)) { ctx->err = "jump set check failed on the device"; return map_status(h[n + i]); }
    }
    return AB200_OK;
}
const ab200::KeyLayout* ab200_ctx_layout(const ab200_ctx* ctx) { return ctx && ctx->have_layout ? &ctx->K : 0; }
int ab200_ctx_swap_system(ab200_ctx* ctx, HostSystem* other) { std::swap(ctx->sys, *other); return upload_system(ctx, true); }
void ab200_ctx_set_error(ab200_ctx* ctx, const char* msg) { if (ctx) ctx->err = msg; }
std::vector<uint64_t>& ab200_cellset_keys(ab200_cellset* cs) { return cs->keys; }
ab200_cellset* ab200_cellset_new() { return new ab200_cellset(); }
