// ab200_kernels.cu -- sm_100a kernels for the grid-reachability lap:
//   K_evolve    one frontier cell per lane; the whole working-set loop of ImageSetHybridEvolver::_evolution
//               (ab200_evolve.cuh) in directed-rounding FP64. Per-cell scratch arenas live in HBM, interleaved
//               word-by-word across the 32 lanes of a warp so lock-step model walks hit one 256-byte line.
//   K_raster    per cell: outer_approximation of its reach / final boxes (ab200_grid.cuh) streamed straight into an
//               open-addressing hash set of packed cell keys (dedupes the ~100x overlap between neighbouring cells
//               before anything is sorted).
//   K_compact / radix sort: unique keys -> sorted array, the form the host cell sets merge.
// The op tapes of the system sit in __constant__ memory (SystemDesc), as every lane reads the same op at the same
// time.
#define AB_COUNT_FLOPS 1
#include <cuda_runtime.h>
#include <stdio.h>
#include "ab200_evolve.cuh"
#include "ab200_jump.cuh"
#include "ab200_kat.cuh"
#include "ab200_grid.cuh"
#include "ab200_device.h"

namespace ab200 { inline namespace AB_VARIANT {

#include "ab200_evolve_kernel.inc"

// ---- flow bounds facade (IntegratorInterface::flow_bounds parity, include/integrator_interface.h:45-62) -----------
__global__ void K_flow_bounds(int n_cells, const int32_t* cell_loc, const double* cell_box, double hmax,
                              double* h_out, double* b_out, int32_t* status) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= n_cells) return;
    Ctx c; c.A.base = 0; c.A.top = 0; c.A.limit = 0; c.A.peak = 0; c.rnd = RN; c.status = ST_OK; c.sm = 0; c.binom = 0; c.sm_dirty = 0; c.flops = 0.0;
    const LocDesc& L = c_sys.loc[cell_loc[cell]];
    Ivl r[AB_MAXN], b[AB_MAXN];
    for (int i = 0; i < L.n; ++i) r[i] = ivl(cell_box[(size_t)cell * 2 * AB_MAXN + 2 * i], cell_box[(size_t)cell * 2 * AB_MAXN + 2 * i + 1]);
    double h = flow_bounds(c, c_sys.ops, L.dyn, L.n, r, hmax, b);
    h_out[cell] = h;
    for (int i = 0; i < L.n; ++i) { b_out[(size_t)cell * 2 * AB_MAXN + 2 * i] = b[i].l; b_out[(size_t)cell * 2 * AB_MAXN + 2 * i + 1] = b[i].u; }
    status[cell] = c.status;
}

// IntegratorBase::flow_bounds with its own constants, one thread per box (include/integrator_interface.h:45-50)
__global__ void K_integrator_flow_bounds(int n_cells, const int32_t* cell_loc, const double* cell_box, double hmax,
                                         double* h_out, double* b_out, int32_t* status) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= n_cells) return;
    Ctx c; c.A.base = 0; c.A.top = 0; c.A.limit = 0; c.A.peak = 0; c.rnd = RN; c.status = ST_OK; c.sm = 0; c.binom = 0; c.sm_dirty = 0; c.flops = 0.0;
    const LocDesc& L = c_sys.loc[cell_loc[cell]];
    Ivl r[AB_MAXN], b[AB_MAXN];
    for (int i = 0; i < L.n; ++i) r[i] = ivl(cell_box[(size_t)cell * 2 * AB_MAXN + 2 * i], cell_box[(size_t)cell * 2 * AB_MAXN + 2 * i + 1]);
    double h = flow_bounds(c, c_sys.ops, L.dyn, L.n, r, hmax, b, true);
    h_out[cell] = h;
    for (int i = 0; i < L.n; ++i) { b_out[(size_t)cell * 2 * AB_MAXN + 2 * i] = b[i].l; b_out[(size_t)cell * 2 * AB_MAXN + 2 * i + 1] = b[i].u; }
    status[cell] = c.status;
}
// TaylorIntegrator::flow for one box, one warp: the n flow models are written to `out` as
// [n_terms, error bits, then n_terms x (key word, coefficient bits)] per model; out[0] = h, out[1] = status, out[2] = words used.
__global__ void __launch_bounds__(32, 1)
K_integrator_flow(int loc, const double* __restrict__ box, double hmax, int order, uint64_t* __restrict__ arena, int arena_words,
                  const uint32_t* __restrict__ binom, uint64_t* __restrict__ out, int out_words) {
    extern __shared__ uint64_t ab_smem[];
    Ctx c; ctx_init(c, arena, arena_words, ab_smem, binom);
    ctx_stage_tables(c);
    const LocDesc& L = c_sys.loc[loc];
    const int n = L.n;
    Ivl dx[AB_MAXN], dom[AB_MAXN + 1]; TM phi[AB_MAXN];
    for (int i = 0; i < n; ++i) dx[i] = ivl(box[2 * i], box[2 * i + 1]);
    const double h = integrator_flow(c, c_sys.ops, L.dyn, n, dx, hmax, order, phi, dom);
    wsync();
    if (wlane() == 0) {
        int w = 3;
        for (int i = 0; i < n && c.status == ST_OK; ++i) {
            if (w + 2 + phi[i].n * AB_TW > out_words) { c.status = ST_OUTPUT; break; }
            out[w++] = (uint64_t)phi[i].n; out[w++] = d2u(phi[i].e);
            for (int t = 0; t < phi[i].n * AB_TW; ++t) out[w++] = c.A.base[phi[i].off + t];
        }
        out[0] = d2u(h); out[1] = (uint64_t)c.status; out[2] = (uint64_t)w;
    }
}

// ---- primitive level entry points (known-answer tests against the reference's own unit-test vectors) -----------------
// One TaylorModel operation, one warp. in: [op, as, k, flag, s, il, iu, nx, xe, (key, coef) x nx, ny, ye, (key, coef) x ny];
// out: [status, words used, range.l, range.u, n_terms, error bits, (key, coef) x n_terms].
__global__ void __launch_bounds__(32, 1)
K_tm_op(const uint64_t* __restrict__ in, uint64_t* __restrict__ arena, int arena_words, const uint32_t* __restrict__ binom,
        uint64_t* __restrict__ out, int out_words) {
    extern __shared__ uint64_t ab_smem[];
    Ctx c; ctx_init(c, arena, arena_words, ab_smem, binom);
    ctx_stage_tables(c);
    const int op = (int)in[0], as = (int)in[1], k = (int)in[2], flag = (int)in[3];
    const double s = u2d(in[4]); const Ivl iv = ivl(u2d(in[5]), u2d(in[6]));
    int p = 7;
    TM m[2];
    for (int q = 0; q < 2; ++q) {
        const int n = (int)in[p++]; const double e = u2d(in[p++]);
        m[q] = tm_begin(c, as, acc_default());
        for (int t = 0; t < n; ++t) { Key key = key_zero(); key.w[0] = in[p]; tm_append(c, m[q], key, u2d(in[p + 1])); p += 2; }
        m[q].e = e;
        wsync();
    }
    Ivl rng = ivl(0.0, 0.0);
    TM r = kat_tm_op(c, op, m[0], m[1], k, flag, s, iv, &rng);
    wsync();
    if (wlane() == 0) {
        int w = 4;
        if (c.status == ST_OK) {
            if (w + 2 + r.n * AB_TW > out_words) c.status = ST_OUTPUT;
            else { out[w++] = (uint64_t)r.n; out[w++] = d2u(r.e); for (int t = 0; t < r.n * AB_TW; ++t) out[w++] = c.A.base[r.off + t]; }
        }
        out[0] = (uint64_t)c.status; out[1] = (uint64_t)w; out[2] = d2u(rng.l); out[3] = d2u(rng.u);
    }
}
// Interval arithmetic, one thread per operation: in[i] = {op, m, a.l, a.u, b.l, b.u}, out[i] = {r.l, r.u, status}.
__global__ void K_interval_op(int n, const double* __restrict__ in, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int op = (int)in[6 * i], m = (int)in[6 * i + 1];
    const Ivl a = ivl(in[6 * i + 2], in[6 * i + 3]), b = ivl(in[6 * i + 4], in[6 * i + 5]);
    int st = ST_OK;
    const Ivl r = kat_interval_op(op, m, a, b, &st);
    out[3 * i] = r.l; out[3 * i + 1] = r.u; out[3 * i + 2] = (double)st;
}

// ---- grid-side jump sets: the activation test of a batch of cell boxes ----------------------------------------------------
// ForwardDiscreteJumpSetChecker::check (src/set_checker.cpp:171-203) for transition `ev` out of location `loc`: one warp per
// box, persistent one-warp CTAs pulling boxes from a counter. tri[b] = 0 false / 1 true / 2 indeterminate.
__global__ void __launch_bounds__(32, 8)
K_jump_check(int n, int loc, int ev, const double* __restrict__ boxes, uint64_t* __restrict__ arena, int arena_words,
             const uint32_t* __restrict__ binom, int* __restrict__ next, int32_t* __restrict__ tri, int32_t* __restrict__ status,
             int32_t* __restrict__ lie_missing) {
    extern __shared__ uint64_t ab_smem[];
    const int lane = wlane();
    for (;;) {
        int b = 0;
        if (lane == 0) b = atomicAdd(next, 1);
        b = wshfl(b, 0);
        if (b >= n) break;
        Ctx c; ctx_init(c, arena + (size_t)blockIdx.x * (size_t)arena_words, arena_words, ab_smem, binom);
        ctx_stage_tables(c);
        double box[2 * AB_MAXN];
        for (int i = 0; i < 2 * AB_MAXN; ++i) box[i] = boxes[(size_t)b * 2 * AB_MAXN + i];
        int lm = 0;
        const int r = jump_check(c, c_sys, loc, ev, box, &lm);
        wsync();
        if (lane == 0) { tri[b] = r; status[b] = c.status; if (lm) atomicOr(lie_missing, 1); }
    }
}

// ---- rasterise + hash-set insert ------------------------------------------------------------------------------------
#define AB_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ void hash_insert(uint64_t* table, uint64_t mask, uint64_t key, unsigned long long* overflow) {
    uint64_t h = key * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    uint64_t slot = h & mask;
    for (uint64_t probe = 0; probe <= mask; ++probe) {
        uint64_t cur = table[slot];
        if (cur == key) return;
        if (cur == AB_EMPTY_KEY) {
            uint64_t old = atomicCAS((unsigned long long*)&table[slot], (unsigned long long)AB_EMPTY_KEY, (unsigned long long)key);
            if (old == AB_EMPTY_KEY || old == key) return;
        }
        slot = (slot + 1) & mask;
    }
    atomicAdd(overflow, 1ull);
}

__global__ void K_fill(uint64_t* p, size_t n, uint64_t v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stridex = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stridex) p[i] = v;
}

// One lane per cell: walks the cell's box list sequentially (the running root depends on the previous boxes), emits
// the covered lattice cells into the hash set of the list (which = 0 reach, 1 final).
__global__ void K_raster(int n_cells, int max_boxes, const double* __restrict__ boxes, const int32_t* __restrict__ box_loc,
                         const int32_t* __restrict__ n_boxes, KeyLayout K, int n_locs,
                         uint64_t* table, uint64_t mask, unsigned long long* counters /* [0]=overflow [1]=out of window [2]=emitted */) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= n_cells) return;
    const int stride = 2 * AB_MAXN;
    const int nb = n_boxes[cell];
    unsigned long long emitted = 0;
    // a fresh running set per location, as outer_approximation builds one DenotableSet per location
    for (int loc = 0; loc < n_locs; ++loc) {
        const LocDesc& L = c_sys.loc[loc];
        const int dim = L.n;
        RunningSet rs; running_init(rs, dim);
        for (int k = 0; k < nb; ++k) {
            if (box_loc[(size_t)cell * max_boxes + k] != loc) continue;
            const double* bp = boxes + ((size_t)cell * max_boxes + k) * stride;
            Ivl b[AB_GRID_MAXN];
            for (int i = 0; i < dim; ++i) b[i] = ivl(bp[2 * i], bp[2 * i + 1]);
            long long lo[AB_GRID_MAXN], hi[AB_GRID_MAXN], idx[AB_GRID_MAXN];
            if (!raster_box(L.grid_origin, L.grid_length, dim, K.acc, rs, b, lo, hi)) { atomicAdd(&counters[1], 1ull); continue; }
            for (int i = 0; i < dim; ++i) idx[i] = lo[i];
            for (;;) {
                uint64_t key;
                if (key_pack(K, loc, idx, &key)) { hash_insert(table, mask, key, &counters[0]); ++emitted; }
                else atomicAdd(&counters[1], 1ull);
                int d = dim - 1;
                while (d >= 0) { if (++idx[d] <= hi[d]) break; idx[d] = lo[d]; --d; }
                if (d < 0) break;
            }
        }
    }
    atomicAdd(&counters[2], emitted);
}

// count[0] = number of keys; count[1] / count[2] = OR / AND of all keys (preset to 0 / ~0): the bits in which the keys differ are
// OR ^ AND, and the radix sort only has to visit the digits that hold one of them.
__global__ void K_compact(const uint64_t* __restrict__ table, size_t n, uint64_t* __restrict__ out, unsigned long long* count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stridex = (size_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    unsigned long long orb = 0ull, andb = ~0ull;
    const size_t rounds = (n + stridex - 1) / stridex;          // every lane of a warp takes part in every ballot
    for (size_t r = 0; r < rounds; ++r, i += stridex) {
        uint64_t k = (i < n) ? table[i] : AB_EMPTY_KEY;
        bool live = (k != AB_EMPTY_KEY);
        unsigned m = __ballot_sync(0xFFFFFFFFu, live);
        // warp-aggregated append
        unsigned long long base = 0;
        if (m) {
            int leader = __ffs(m) - 1;
            if (lane == leader) base = atomicAdd(count, (unsigned long long)__popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (live) { out[base + __popc(m & ((1u << lane) - 1))] = k; orb |= k; andb &= k; }
        }
    }
    for (int d = 16; d > 0; d >>= 1) { orb |= __shfl_xor_sync(0xFFFFFFFFu, orb, d); andb &= __shfl_xor_sync(0xFFFFFFFFu, andb, d); }
    if (lane == 0 && orb != 0ull) { atomicOr(count + 1, orb); atomicAnd(count + 2, andb); }
}

// ---- LSD radix sort (8-bit digits), stable, one warp per tile -------------------------------------------------------
#define AB_SORT_TILE 2048
__global__ void K_sort_hist(const uint64_t* __restrict__ in, size_t n, int shift, unsigned* __restrict__ hist /* [256][nblocks] */, int nblocks) {
    __shared__ unsigned h[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    size_t start = (size_t)blockIdx.x * AB_SORT_TILE, end = start + AB_SORT_TILE; if (end > n) end = n;
    for (size_t i = start + threadIdx.x; i < end; i += blockDim.x) atomicAdd(&h[(in[i] >> shift) & 0xFF], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[(size_t)i * nblocks + blockIdx.x] = h[i];
}
// exclusive scan of hist in (digit-major, block-minor) order; single block.
__global__ void K_sort_scan(unsigned* hist, size_t total) {
    __shared__ unsigned carry;
    __shared__ unsigned buf[1024];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (size_t base = 0; base < total; base += 1024) {
        size_t i = base + threadIdx.x;
        unsigned v = (i < total) ? hist[i] : 0;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            unsigned t = (threadIdx.x >= off) ? buf[threadIdx.x - off] : 0;
            __syncthreads();
            buf[threadIdx.x] += t;
            __syncthreads();
        }
        unsigned incl = buf[threadIdx.x];
        if (i < total) hist[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
}
__global__ void K_sort_scatter(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, size_t n, int shift,
                               const unsigned* __restrict__ hist, int nblocks) {
    __shared__ unsigned offs[256];
    for (int i = threadIdx.x; i < 256; i += 32) offs[i] = hist[(size_t)i * nblocks + blockIdx.x];
    __syncwarp();
    size_t start = (size_t)blockIdx.x * AB_SORT_TILE, end = start + AB_SORT_TILE; if (end > n) end = n;
    const int lane = threadIdx.x;
    for (size_t base = start; base < end; base += 32) {
        size_t i = base + lane;
        bool live = i < end;
        uint64_t k = live ? in[i] : 0;
        unsigned d = live ? (unsigned)((k >> shift) & 0xFF) : 0x100u + lane;   // distinct dummy digits for dead lanes
        unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
        unsigned rank = __popc(peers & ((1u << lane) - 1));
        unsigned pos = 0;
        if (live) pos = offs[d] + rank;
        __syncwarp();
        if (live && rank == 0) offs[d] += __popc(peers);
        __syncwarp();
        if (live) out[pos] = k;
    }
}

// ---- merge of gathered key arrays (multi-GPU exchange): flag the first occurrence of every key of a sorted array, scan, scatter
__global__ void K_unique_flag(const uint64_t* __restrict__ in, size_t n, unsigned* __restrict__ flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stridex = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stridex) flag[i] = (in[i] != AB_EMPTY_KEY && (i == 0 || in[i] != in[i - 1])) ? 1u : 0u;
}
__global__ void K_unique_scatter(const uint64_t* __restrict__ in, size_t n, const unsigned* __restrict__ pos, uint64_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stridex = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stridex) if (in[i] != AB_EMPTY_KEY && (i == 0 || in[i] != in[i - 1])) out[pos[i]] = in[i];
}

// ---- FP64 pipe peak probe (reported beside the roofline) -------------------------------------------------
__global__ void K_fp64_peak(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 0.1, a2 = a0 + 0.2, a3 = a0 + 0.3, a4 = a0 + 0.4, a5 = a0 + 0.5, a6 = a0 + 0.6, a7 = a0 + 0.7;
    const double m = 1.0000001, b = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = __dadd_ru(__dmul_rd(a0, m), b); a1 = __dadd_ru(__dmul_rd(a1, m), b); a2 = __dadd_ru(__dmul_rd(a2, m), b); a3 = __dadd_ru(__dmul_rd(a3, m), b);
        a4 = __dadd_ru(__dmul_rd(a4, m), b); a5 = __dadd_ru(__dmul_rd(a5, m), b); a6 = __dadd_ru(__dmul_rd(a6, m), b); a7 = __dadd_ru(__dmul_rd(a7, m), b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// The same probe with fused multiply-adds (the nominal FP64 peak counts an FMA as two operations; the directed-rounding
// arithmetic of this library cannot use FMAs -- every product and every sum is rounded separately).
__global__ void K_fp64_peak_fma(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 0.1, a2 = a0 + 0.2, a3 = a0 + 0.3, a4 = a0 + 0.4, a5 = a0 + 0.5, a6 = a0 + 0.6, a7 = a0 + 0.7;
    const double m = 1.0000001, b = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = __fma_rn(a0, m, b); a1 = __fma_rn(a1, m, b); a2 = __fma_rn(a2, m, b); a3 = __fma_rn(a3, m, b);
        a4 = __fma_rn(a4, m, b); a5 = __fma_rn(a5, m, b); a6 = __fma_rn(a6, m, b); a7 = __fma_rn(a7, m, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ---- host-callable launchers ------------------------------------------------------------------------------------------
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { snprintf(errbuf, errlen, "%s: %s", #x, cudaGetErrorString(e_)); return (int)e_; } } while (0)

int dev_upload_system(const SystemDesc* s, char* errbuf, int errlen) { return dev_upload_system_v(s, errbuf, errlen); }
int dev_set_stack(size_t bytes, char* errbuf, int errlen) {
    CK(cudaDeviceSetLimit(cudaLimitStackSize, bytes));
    return 0;
}
int dev_launch_flow_bounds(cudaStream_t st, int n_cells, const int32_t* cell_loc, const double* cell_box, double hmax,
                           double* h_out, double* b_out, int32_t* status, char* errbuf, int errlen) {
    if (n_cells == 0) return 0;
    K_flow_bounds<<<(n_cells + 127) / 128, 128, 0, st>>>(n_cells, cell_loc, cell_box, hmax, h_out, b_out, status);
    CK(cudaGetLastError());
    return 0;
}
int dev_launch_integrator_flow_bounds(cudaStream_t st, int n_cells, const int32_t* cell_loc, const double* cell_box, double hmax,
                                      double* h_out, double* b_out, int32_t* status, char* errbuf, int errlen) {
    if (n_cells == 0) return 0;
    K_integrator_flow_bounds<<<(n_cells + 127) / 128, 128, 0, st>>>(n_cells, cell_loc, cell_box, hmax, h_out, b_out, status);
    CK(cudaGetLastError());
    return 0;
}
int dev_launch_integrator_flow(cudaStream_t st, int loc, const double* box, double hmax, int order, uint64_t* arena, int arena_words,
                               const uint32_t* binom, uint64_t* out, int out_words, char* errbuf, int errlen) {
    K_integrator_flow<<<1, 32, AB_SM_WORDS * sizeof(uint64_t), st>>>(loc, box, hmax, order, arena, arena_words, binom, out, out_words);
    CK(cudaGetLastError());
    return 0;
}
int dev_launch_jump_check(cudaStream_t st, int n, int blocks, int loc, int ev, const double* boxes, uint64_t* arena, int arena_words,
                          const uint32_t* binom, int* next, int32_t* tri, int32_t* status, int32_t* lie_missing, char* errbuf, int errlen) {
    if (n == 0) return 0;
    K_jump_check<<<blocks, 32, AB_SM_WORDS * sizeof(uint64_t), st>>>(n, loc, ev, boxes, arena, arena_words, binom, next, tri, status, lie_missing);
    CK(cudaGetLastError());
    return 0;
}
int dev_launch_tm_op(cudaStream_t st, const uint64_t* in, uint64_t* arena, int arena_words, const uint32_t* binom, uint64_t* out, int out_words,
                     char* errbuf, int errlen) {
    K_tm_op<<<1, 32, AB_SM_WORDS * sizeof(uint64_t), st>>>(in, arena, arena_words, binom, out, out_words);
    CK(cudaGetLastError());
    return 0;
}
int dev_launch_interval_op(cudaStream_t st, int n, const double* in, double* out, char* errbuf, int errlen) {
    if (n == 0) return 0;
    K_interval_op<<<(n + 127) / 128, 128, 0, st>>>(n, in, out);
    CK(cudaGetLastError());
    return 0;
}
int dev_launch_fill(cudaStream_t st, uint64_t* p, size_t n, uint64_t v, char* errbuf, int errlen) {
    if (n == 0) return 0;
    K_fill<<<148 * 8, 256, 0, st>>>(p, n, v);
    CK(cudaGetLastError());
    return 0;
}
int dev_launch_raster(cudaStream_t st, int n_cells, int max_boxes, const double* boxes, const int32_t* box_loc, const int32_t* n_boxes,
                      const KeyLayout& K, int n_locs, uint64_t* table, uint64_t mask, unsigned long long* counters, char* errbuf, int errlen) {
    if (n_cells == 0) return 0;
    K_raster<<<(n_cells + 127) / 128, 128, 0, st>>>(n_cells, max_boxes, boxes, box_loc, n_boxes, K, n_locs, table, mask, counters);
    CK(cudaGetLastError());
    return 0;
}
int dev_launch_compact(cudaStream_t st, const uint64_t* table, size_t n, uint64_t* out, unsigned long long* count, char* errbuf, int errlen) {
    K_compact<<<148 * 8, 256, 0, st>>>(table, n, out, count);
    CK(cudaGetLastError());
    return 0;
}
// Sorts keys[0..n) ascending on the low `bits` bits; tmp and hist are scratch (hist: 256*ceil(n/TILE) unsigned).
// `varying`: the bits in which the keys differ (all ones if unknown); digits without such a bit are skipped.
int dev_radix_sort(cudaStream_t st, uint64_t* keys, uint64_t* tmp, size_t n, int bits, unsigned long long varying, unsigned* hist, int* result_in_tmp,
                   char* errbuf, int errlen) {
    *result_in_tmp = 0;
    if (n <= 1) return 0;
    int nblocks = (int)((n + AB_SORT_TILE - 1) / AB_SORT_TILE);
    uint64_t* src = keys; uint64_t* dst = tmp;
    for (int shift = 0; shift < bits; shift += 8) {
        if (((varying >> shift) & 0xFFull) == 0ull) continue;
        K_sort_hist<<<nblocks, 256, 0, st>>>(src, n, shift, hist, nblocks);
        K_sort_scan<<<1, 1024, 0, st>>>(hist, (size_t)256 * nblocks);
        K_sort_scatter<<<nblocks, 32, 0, st>>>(src, dst, n, shift, hist, nblocks);
        CK(cudaGetLastError());
        uint64_t* t = src; src = dst; dst = t;
        *result_in_tmp ^= 1;
    }
    return 0;
}
int dev_sort_tile() { return AB_SORT_TILE; }
// Sorted `in` (n keys, AB_EMPTY_KEY padding allowed) -> duplicate-free `out`; pos: n+1 unsigned of scratch. The number of keys
// kept is pos[n-1] + (flag of the last key), read back by the caller from pos_last / flag_last.
int dev_unique_sorted(cudaStream_t st, const uint64_t* in, size_t n, unsigned* pos, uint64_t* out, char* errbuf, int errlen) {
    if (n == 0) return 0;
    K_unique_flag<<<148 * 4, 256, 0, st>>>(in, n, pos);
    K_sort_scan<<<1, 1024, 0, st>>>(pos, n);
    K_unique_scatter<<<148 * 4, 256, 0, st>>>(in, n, pos, out);
    CK(cudaGetLastError());
    return 0;
}
int dev_fp64_peak(cudaStream_t st, double* scratch, int blocks, int threads, int iters, int fma, char* errbuf, int errlen) {
    if (fma) K_fp64_peak_fma<<<blocks, threads, 0, st>>>(scratch, iters);
    else K_fp64_peak<<<blocks, threads, 0, st>>>(scratch, iters);
    CK(cudaGetLastError());
    return 0;
}

} } // namespace ab200
