// ab200_tm.cuh -- TaylorModel primitives, one warp per frontier cell.
//
// A TaylorModel is a sparse polynomial on [-1,1]^as with a uniform error bound (reference:
// include/taylor_model.h:170-190, include/expansion.h, include/multi_index.h). Here a model is a warp-uniform handle
// {offset,count,as,error,accuracy} into the cell's scratch arena of 64-bit words (global memory; a few thousand cells
// are in flight, so the arenas stay resident in L2); terms are [key words | coefficient] records kept in the
// reference's graded order (include/multi_index.h:346-358). Loops over terms are strided over the lanes of the warp.
//
// Sums of models -- the x-term streams of a product (_mul4), the batched additions of _compose2 and _split1, _add2,
// intersection -- go through ONE mechanism, a sparse accumulator staged in shared memory (AccTab): a structure-of-arrays
// list of (MultiIndex key, coefficient) plus a hash index from key to list position. A stream is added by all lanes at
// once (the keys of one stream are distinct, so no two lanes touch the same entry); streams are added one after another,
// which is the reference's fold order for every coefficient. The accumulator is finally written to the arena in key
// order. This replaces the reference's re-merge of the whole accumulator per x-term (src/taylor_model.cpp:772-804).
//
// Error bounds: every rounding gap up(u+ml) of a coincident term is either 0 or a power of two (u and -ml are equal or
// adjacent doubles), so the gaps of a product are recorded as 16-bit exponent codes per (x-term, y-term) and summed
// afterwards by one lane per x-term, upward and in y order -- exactly the reference's te of each _add2 -- while the lanes
// run side by side. Discarded products are flagged in the same code and their truncation error |x||y| is recomputed.
//
// MultiIndex packing: AB_KW 64-bit words per key. Word 0 = degree byte (bits 56..63) followed by the exponents of
// variables 0..6, most significant first; word 1 = exponents of variables 7..14. Adding two keys is a plain integer add
// per word; the reference order is "degree ascending, then the key whose first differing exponent is LARGER comes
// first", i.e. ascending (w0 ^ 0x00FF..FF) then descending w1.
//
// Every accumulator keeps the reference's association order, so results are bit-identical to the CPU library; see
// DESIGN.md "Numerical contract".
#pragma once
#include "ab200_warp.cuh"

#ifndef AB_KW
#define AB_KW 1
#endif

namespace ab200 { inline namespace AB_VARIANT {

#define AB_TW (AB_KW + 1)   // words per term
#define AB_MAXARGS (AB_KW == 1 ? 7 : 15)

struct Key {
    uint64_t w[AB_KW];
};

AB_HD Key key_zero() { Key k; for (int i = 0; i < AB_KW; ++i) k.w[i] = 0; return k; }
AB_HD int key_deg(const Key& k) { return (int)(k.w[0] >> 56); }
AB_HD int key_get(const Key& k, int j) {
    if (j < 7) return (int)((k.w[0] >> (8 * (6 - j))) & 0xFF);
#if AB_KW > 1
    return (int)((k.w[1] >> (8 * (14 - j))) & 0xFF);
#else
    return 0;
#endif
}
// add d (possibly negative) to exponent j, maintaining the degree byte
AB_HD void key_inc(Key& k, int j, int d) {
    if (j < 7) k.w[0] += (uint64_t)((int64_t)d) * ((1ull << (8 * (6 - j))) + (1ull << 56));
#if AB_KW > 1
    else { k.w[1] += (uint64_t)((int64_t)d) << (8 * (14 - j)); k.w[0] += (uint64_t)((int64_t)d) << 56; }
#endif
}
// the key with exponent j set to e (e >= 0)
AB_HD Key key_with(const Key& a, int j, int e) { Key r = a; key_inc(r, j, e - key_get(a, j)); return r; }
// per-byte maximum of two packed words (exponents are below 128, so the top bit of every byte is free)
AB_HD uint64_t bytes_max(uint64_t x, uint64_t y) {
    const uint64_t H = 0x8080808080808080ull;
    const uint64_t ge = (((x | H) - y) & H) >> 7;      // 1 in every byte where x >= y
    const uint64_t m = ge * 0xFFull;
    return (x & m) | (y & ~m);
}
AB_HD Key key_unit(int j) { Key k = key_zero(); key_inc(k, j, 1); return k; }
AB_HD Key key_add(const Key& a, const Key& b) { Key r; for (int i = 0; i < AB_KW; ++i) r.w[i] = a.w[i] + b.w[i]; return r; }
AB_HD bool key_eq(const Key& a, const Key& b) {
    bool e = a.w[0] == b.w[0];
#if AB_KW > 1
    e = e && a.w[1] == b.w[1];
#endif
    return e;
}
AB_HD bool key_less(const Key& a, const Key& b) {   // include/multi_index.h:346-358
    const uint64_t M = 0x00FFFFFFFFFFFFFFull;
    uint64_t sa = a.w[0] ^ M, sb = b.w[0] ^ M;
    if (sa != sb) return sa < sb;
#if AB_KW > 1
    return a.w[1] > b.w[1];
#else
    return false;
#endif
}
// remove variable k (shift later variables down); degree byte reduced by the removed exponent
AB_HD Key key_remove(const Key& a, int k, int as) {
    Key r = key_zero();
    int jj = 0;
    #pragma unroll 1
    for (int j = 0; j < as; ++j) { if (j == k) continue; int e = key_get(a, j); if (e) key_inc(r, jj, e); ++jj; }
    return r;
}

struct Acc {          // TaylorModel::Accuracy (include/taylor_model.h:178-190)
    double sweep;
    int maxdeg;
};
AB_HD Acc acc_default() { Acc a; a.sweep = 1e-18; a.maxdeg = 16; return a; }   // src/taylor_model.cpp:46-47
AB_HD bool acc_discard(const Acc& a, const Key& k, double v) { return dabs(v) < a.sweep || key_deg(k) > a.maxdeg; }

struct Arena {
    uint64_t* base;   // the cell's scratch words (global memory)
    int top;
    int limit;
    int peak;
};

// Every field of Ctx is warp-uniform (all lanes of the cell's warp hold the same value at the same point of the program), so
// the kernels keep ONE copy per warp in shared memory instead of 32 thread-local copies on the stack: the noinline primitives
// take it by reference, and a stack copy costs a 256-byte local-memory transaction per field access. The lane index is
// therefore not a field: use wlane().
struct Ctx {
    Arena A;
    int rnd;       // ambient rounding mode left behind by the last routine (mirrors MXCSR in the reference)
    int status;    // sticky first error
    uint64_t* sm;  // the warp's shared-memory staging area (AB_SM_WORDS words); null in host-side callers
    const uint32_t* binom;   // binomial table for monomial ranks (binom_fill)
    int sm_dirty;            // the hash accumulator has written over the dense accumulator's area since it was last emptied
    double flops;            // algorithmic FP64 operations of the reference algorithm so far (AB_FLOPS)
    // phase alignment of the warps of a CTA (K_evolve): shared progress words, warps in the CTA, this warp, loop iteration
    int* al_phase; int al_n; int al_warp; int al_iter;
    int libm_sensitive;      // rec / sqrt calls whose result could change with the last bits of log / pow (ab200_series.cuh)
    // tm_compose: power-table base and accuracy per inner model; arrays indexed at run time, kept here (shared memory in the
    // kernels) rather than on the stack, where every lookup in the term x factor loop would be a local-memory transaction
    int cmp_tab[AB_MAXARGS + 1];
    Acc cmp_acc[AB_MAXARGS + 1];
};
#define AB_BINOM_N 64
#define AB_BINOM_K 24
// C(n,k) for n < AB_BINOM_N, k < AB_BINOM_K, saturated; filled on the host and handed to the kernels in global memory.
inline void binom_fill(uint32_t* t) {
    #pragma unroll 1
    for (int n = 0; n < AB_BINOM_N; ++n) for (int k = 0; k < AB_BINOM_K; ++k) {
        unsigned long long v = (k == 0) ? 1 : (n == 0 ? 0 : (unsigned long long)t[(n - 1) * AB_BINOM_K + k - 1] + t[(n - 1) * AB_BINOM_K + k]);
        t[n * AB_BINOM_K + k] = v > 0x7FFFFFFFull ? 0x7FFFFFFFu : (uint32_t)v; }
}
struct BinomTab { uint32_t t[AB_BINOM_N * AB_BINOM_K]; BinomTab() { binom_fill(t); } };
inline const uint32_t* binom_host() { static BinomTab b; return b.t; }
AB_HD void ctx_init(Ctx& c, uint64_t* arena, int limit, uint64_t* sm, const uint32_t* binom) {
    c.A.base = arena; c.A.top = 0; c.A.limit = limit; c.A.peak = 0; c.rnd = RN; c.status = ST_OK; c.sm = sm; c.binom = binom; c.sm_dirty = 1; c.flops = 0.0;
    c.al_phase = 0; c.al_n = 1; c.al_warp = 0; c.al_iter = 0; c.libm_sensitive = 0;
}
// host-side callers (no warp, no shared area)
inline void ctx_init_host(Ctx& c, uint64_t* arena, int limit) {
    c.A.base = arena; c.A.top = 0; c.A.limit = limit; c.A.peak = 0; c.rnd = RN; c.status = ST_OK; c.sm = 0; c.binom = binom_host(); c.sm_dirty = 1; c.flops = 0.0;
    c.al_phase = 0; c.al_n = 1; c.al_warp = 0; c.al_iter = 0; c.libm_sensitive = 0;
}

// Phase alignment (device, CTAs of several warps): the warps of a CTA evolve different cells through the same sequence of
// primitives; meeting at the main stages of a step keeps them in the same few kilobytes of code, which is what the SM's
// instruction cache can hold. A warp publishes (loop iteration, stage) and waits until no warp of its CTA is behind that
// stage. A warp that skips a stage publishes a later one, and the end of the loop iteration is stage 63, so the warp
// that is furthest behind never waits and nobody waits for ever. Pure scheduling: no arithmetic depends on it.
#ifndef AB_ALIGN_SPINS
#define AB_ALIGN_SPINS (1 << 18)
#endif
AB_HD void cta_publish(Ctx& c, int stage) {
#if defined(__CUDA_ARCH__)
    if (c.al_n > 1 && wlane() == 0) ((volatile int*)c.al_phase)[c.al_warp] = (c.al_iter << 6) | stage;
#endif
}
AB_HDN void cta_align(Ctx& c, int stage) {
#if defined(__CUDA_ARCH__)
    if (c.al_n <= 1) return;
    if (stage > 62) stage = 62;
    const int my = (c.al_iter << 6) | stage;
    volatile int* ph = (volatile int*)c.al_phase;
    __syncwarp();
    if (wlane() == 0) ph[c.al_warp] = my;
    int spins = 0;
    for (;;) {
        int lo = my;
        #pragma unroll 1
        for (int w = 0; w < c.al_n; ++w) { const int v = ph[w]; if (v < lo) lo = v; }
        if (lo >= my || ++spins > AB_ALIGN_SPINS) break;
        __nanosleep(40);
    }
    __syncwarp();
#endif
}
AB_HD uint64_t& aw(Ctx& c, int i) { return c.A.base[i]; }
AB_HD void fail(Ctx& c, int st) { if (c.status == ST_OK) c.status = st; }
#ifdef AB_MUL_HIST
extern long long g_calls[16], g_words[16];   // development counters (tests/tools/hostdbg.cpp)
#define AB_CALL(id, w) do { if (wlane() == 0) { g_calls[id]++; g_words[id] += (w); } } while (0)
// Algorithmic FP64 operations of the REFERENCE algorithm by the per-primitive formulas of SURVEY.md 8(d), evaluated on
// the actual term counts (host development builds only; the figure bench.py quotes comes from these counters).
extern double g_flops;
#define AB_FLOPS(n) do { if (wlane() == 0) g_flops += (double)(n); } while (0)
#elif defined(AB_COUNT_FLOPS)
// Device builds: the same count kept per cell (warp-uniform, one FP64 add per primitive call) and returned with the lap's
// status, so that the FP64 roofline of bench.py is measured on the run itself.
#define AB_CALL(id, w)
#define AB_FLOPS(n) do { c.flops += (double)(n); } while (0)
#else
#define AB_CALL(id, w)
#define AB_FLOPS(n)
#endif
#define AB_SECTION(id)

struct TM {
    int off;      // first word in the arena
    int n;        // number of terms
    int as;       // argument size
    double e;     // uniform error
    Acc acc;
};

AB_HD Key bkey(const uint64_t* base, int w) { Key k; for (int j = 0; j < AB_KW; ++j) k.w[j] = base[w + j]; return k; }
AB_HD double bval(const uint64_t* base, int w) { return u2d(base[w + AB_KW]); }
AB_HD void bput(uint64_t* base, int w, const Key& k, double v) { for (int j = 0; j < AB_KW; ++j) base[w + j] = k.w[j]; base[w + AB_KW] = d2u(v); }

AB_HD Key tm_key(Ctx& c, const TM& x, int i) { return bkey(c.A.base, x.off + i * AB_TW); }
AB_HD double tm_val(Ctx& c, const TM& x, int i) { return bval(c.A.base, x.off + i * AB_TW); }
AB_HD void tm_set_val(Ctx& c, const TM& x, int i, double v) { c.A.base[x.off + i * AB_TW + AB_KW] = d2u(v); }
AB_HD void tm_set_key(Ctx& c, const TM& x, int i, const Key& k) { for (int j = 0; j < AB_KW; ++j) c.A.base[x.off + i * AB_TW + j] = k.w[j]; }
AB_HD int tm_end(const TM& x) { return x.off + x.n * AB_TW; }

AB_HD void arena_mark_peak(Ctx& c) { if (c.A.top > c.A.peak) c.A.peak = c.A.top; }
// Reserve room for `terms` more terms above `from`; false (and ST_ARENA) if the arena cannot hold them.
AB_HD bool arena_room(Ctx& c, int from, int terms) {
    if ((long long)from + (long long)terms * AB_TW > (long long)c.A.limit) { fail(c, ST_ARENA); return false; }
    return true;
}

// Begin a model at the arena top; append() grows it. Only the most recently begun model may be appended to.
// tm_append is a warp-uniform operation: every lane writes the same term (one transaction), so the term is visible to
// each lane without a barrier.
AB_HD TM tm_begin(Ctx& c, int as, Acc acc) { TM r; r.off = c.A.top; r.n = 0; r.as = as; r.e = 0.0; r.acc = acc; return r; }
AB_HD void tm_append(Ctx& c, TM& r, const Key& k, double v) {
    int o = r.off + r.n * AB_TW;
    if (o + AB_TW > c.A.limit) { fail(c, ST_ARENA); return; }
    bput(c.A.base, o, k, v);
    r.n++;
    c.A.top = o + AB_TW;
    arena_mark_peak(c);
}
// Move model x (anywhere at or above dst) down to word offset dst and set the arena top just past it.
AB_HDN void tm_slide(Ctx& c, TM& x, int dst) { AB_CALL(2, x.off == dst ? 0 : x.n);
    if (x.off != dst) {
        wsync();
        const int nw = x.n * AB_TW, src = x.off, W = AB_W, lane = wlane();
        uint64_t* const base = c.A.base;
        if (src - dst >= nw) { for (int i = lane; i < nw; i += W) base[dst + i] = base[src + i]; }
        else for (int i0 = 0; i0 < nw; i0 += W) {      // overlapping: a chunk is read by all lanes before it is written
            const int i = i0 + lane; uint64_t v = 0;
            if (i < nw) v = base[src + i];
            wsync();
            if (i < nw) base[dst + i] = v;
        }
        wsync();
        x.off = dst;
    }
    c.A.top = dst + x.n * AB_TW;
}
AB_HDN TM tm_copy(Ctx& c, const TM& x) { AB_CALL(3, x.n);
    wsync();
    TM r = tm_begin(c, x.as, x.acc);
    const int nw = x.n * AB_TW;
    if (!arena_room(c, r.off, x.n)) return r;
    { uint64_t* const base = c.A.base; const int dst = r.off, src = x.off, W = AB_W;
      #pragma unroll 1
      for (int i = wlane(); i < nw; i += W) base[dst + i] = base[src + i]; }
    wsync();
    r.n = x.n; r.e = x.e;
    c.A.top = r.off + nw;
    arena_mark_peak(c);
    return r;
}
AB_HD TM tm_zero(Ctx& c, int as, Acc acc) { return tm_begin(c, as, acc); }
AB_HD int tm_degree(Ctx& c, const TM& x) { return x.n == 0 ? 0 : key_deg(tm_key(c, x, x.n - 1)); }   // taylor_model.h:361

// ---- in-place scalar operations -------------------------------------------------------------------------------

// _scal(TaylorModel&, Float) src/taylor_model.cpp:241-273
AB_HDN void tm_scal(Ctx& c, TM& r, double s) { AB_FLOPS(4 * r.n + 3); AB_CALL(4, r.n);
    if (s == 0) { r.n = 0; r.e = 0; return; }     // (arena words stay allocated; harmless)
    if (s == 1) return;
    const int W = AB_W, lane = wlane(), n = r.n;
    wsync();
    if (s == 2 || s == 0.5) {
        const int m = c.rnd;
        #pragma unroll 1
        for (int i = lane; i < n; i += W) tm_set_val(c, r, i, mul_m(m, tm_val(c, r, i), s));
        wsync();
        r.e = mul_m(c.rnd, r.e, dabs(s));
        return;
    }
    double te = 0; const double ms = -s;
    #pragma unroll 1
    for (int i0 = 0; i0 < n; i0 += W) {
        const int i = i0 + lane; double g = 0.0;
        if (i < n) {
            const double rv = tm_val(c, r, i);
            g = add_ru(mul_ru(rv, s), mul_ru(rv, ms));
            tm_set_val(c, r, i, mul_rn(rv, s));
        }
        te = chain_ru(te, g, (n - i0 < W) ? n - i0 : W);
    }
    wsync();
    r.e = mul_ru(r.e, dabs(s));
    r.e = add_ru(r.e, te / 2);
    c.rnd = RN;
}
// _scal(TaylorModel&, Interval) src/taylor_model.cpp:276-314
AB_HDN void tm_scal(Ctx& c, TM& r, Ivl s) { AB_FLOPS(5 * r.n + 4); AB_CALL(5, r.n);
    const int W = AB_W, lane = wlane(), n = r.n;
    wsync();
    double te = 0; const double cu = s.u, mcl = -s.l;
    const double m = add_rn(s.u, s.l) / 2;
    #pragma unroll 1
    for (int i0 = 0; i0 < n; i0 += W) {
        const int i = i0 + lane; double g = 0.0;
        if (i < n) {
            const double rv = tm_val(c, r, i); double u, ml;
            if (rv >= 0) { u = mul_ru(rv, cu); ml = mul_ru(rv, mcl); }
            else { double mrv = -rv; u = mul_ru(mrv, mcl); ml = mul_ru(mrv, cu); }
            g = add_ru(u, ml);
            tm_set_val(c, r, i, mul_rn(rv, m));
        }
        te = chain_ru(te, g, (n - i0 < W) ? n - i0 : W);
    }
    wsync();
    r.e = mul_ru(r.e, ivl_mag(s));
    r.e = add_ru(r.e, te / 2);
    c.rnd = RN;
}
// Insert a constant term in front of a model that is the LAST thing on the arena.
AB_HDN void tm_prepend_top(Ctx& c, TM& r, double v) {
    const int nw = r.n * AB_TW, W = AB_W, lane = wlane();
    if (!arena_room(c, r.off, r.n + 1)) return;
    uint64_t* const base = c.A.base;
    wsync();
    #pragma unroll 1
    for (int i0 = nw > 0 ? ((nw - 1) / W) * W : -1; i0 >= 0; i0 -= W) {     // shift up, last chunk first
        const int i = i0 + lane; uint64_t t = 0;
        if (i < nw) t = base[r.off + i];
        wsync();
        if (i < nw) base[r.off + AB_TW + i] = t;
    }
    wsync();
    bput(base, r.off, key_zero(), v);
    r.n++;
    c.A.top = r.off + r.n * AB_TW;
    arena_mark_peak(c);
}
// _acc(TaylorModel&, Float) src/taylor_model.cpp:370-392. r must be the top model on the arena.
AB_HDN void tm_acc(Ctx& c, TM& r, double s) { AB_FLOPS(5);
    if (s == 0) return;
    if (r.n == 0 || key_deg(tm_key(c, r, 0)) > 0) { tm_prepend_top(c, r, s); return; }
    const double rv = tm_val(c, r, 0);
    wsync();                                   // every lane has read the old constant before it is replaced
    double rvu = add_ru(rv, s), mrvl = sub_ru(-rv, s);
    r.e = add_ru(r.e, add_ru(rvu, mrvl) / 2);
    tm_set_val(c, r, 0, add_rn(rv, s));
    c.rnd = RN;
}
// _acc(TaylorModel&, Interval) src/taylor_model.cpp:397-422. r must be the top model on the arena.
AB_HDN void tm_acc(Ctx& c, TM& r, Ivl s) { AB_FLOPS(6);
    if (s.l == -s.u) { r.e = add_ru(r.e, s.u); c.rnd = RN; return; }
    if (r.n == 0 || key_deg(tm_key(c, r, 0)) > 0) tm_prepend_top(c, r, 0.0);
    const double rv = tm_val(c, r, 0);
    wsync();
    double rvu = add_ru(rv, s.u), mrvl = sub_ru(-rv, s.l);
    r.e = add_ru(r.e, add_ru(rvu, mrvl) / 2);
    double m = add_rn(s.u, s.l) / 2;
    tm_set_val(c, r, 0, add_rn(rv, m));
    c.rnd = RN;
}

// ---- the accumulators ---------------------------------------------------------------------------------------------
// Shared-memory staging area of one warp, in 64-bit words:
//   [ accumulator: AB_SM_ACCW | rank table: AB_SM_BTABW | gap codes: AB_SM_G16 uint16 | staged factor: keys AB_KW x AB_SM_YCAP,
//     coefficients AB_SM_YCAP ]
// The accumulator area holds either the DENSE accumulator (structure-of-arrays keys | coefficients indexed by the monomial's
// rank in the graded order: no search, no insertion, emission is a scan) when the monomial space of the result fits, or the
// hash accumulator (index | keys | coefficients) otherwise. When even that is outgrown the hash accumulator is rebuilt in the
// free part of the arena (same code, generic pointers).
#ifndef AB_SM_ACCW
#define AB_SM_ACCW 1536
#endif
#define AB_SM_BTABW 128
// Dense accumulator: coefficients of AB_SM_DENSE ranks, then their keys. With at most 3 variables a key is the upper half
// of the packed MultiIndex word (degree and three exponent bytes), so AB_SM_DENSE keys take half a word each; with 4..7
// variables the full word is kept and only AB_SM_DENSE/2 ranks are available.
#define AB_SM_DENSE (2 * AB_SM_ACCW / 3)
#define AB_BTAB_STRIDE 36                               // rank-table row: sums of exponents 0..35
#ifndef AB_SM_NIDX
#define AB_SM_NIDX 1024
#endif
#define AB_SM_CAP ((AB_SM_ACCW - AB_SM_NIDX / 2) / (AB_KW + 1))
#ifndef AB_SM_G16
#define AB_SM_G16 2048
#endif
#ifndef AB_SM_YCAP
#define AB_SM_YCAP 128
#endif
#define AB_SM_OFF_IDX 0
#define AB_SM_OFF_KEY (AB_SM_NIDX / 2)
#define AB_SM_OFF_VAL (AB_SM_OFF_KEY + AB_KW * AB_SM_CAP)
#define AB_SM_OFF_BTAB AB_SM_ACCW
#define AB_SM_OFF_G16 (AB_SM_ACCW + AB_SM_BTABW)
#define AB_SM_OFF_YK (AB_SM_OFF_G16 + AB_SM_G16 / 4)
#define AB_SM_OFF_YV (AB_SM_OFF_YK + AB_KW * AB_SM_YCAP)
#define AB_SM_WORDS (AB_SM_OFF_YV + AB_SM_YCAP)

// Shared-memory accesses by 32-bit shared address (device) so that the hot loops use LDS/STS instead of generic loads with
// 64-bit address arithmetic; plain pointers in the host builds.
#if defined(__CUDA_ARCH__)
typedef uint32_t smaddr_t;
AB_HD smaddr_t sm_addr(const void* p) { return (smaddr_t)__cvta_generic_to_shared(p); }
AB_HD uint64_t sm_ld64(smaddr_t a) { unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return (uint64_t)v; }
AB_HD void sm_st64(smaddr_t a, uint64_t v) { asm volatile("st.shared.u64 [%0], %1;" :: "r"(a), "l"((unsigned long long)v) : "memory"); }
AB_HD uint32_t sm_ld32(smaddr_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
AB_HD void sm_st32(smaddr_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
AB_HD uint32_t sm_ld16(smaddr_t a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory"); return (uint32_t)v; }
AB_HD void sm_st16(smaddr_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"((uint16_t)v) : "memory"); }
#else
typedef uintptr_t smaddr_t;
inline smaddr_t sm_addr(const void* p) { return (smaddr_t)p; }
inline uint64_t sm_ld64(smaddr_t a) { return *(const uint64_t*)a; }
inline void sm_st64(smaddr_t a, uint64_t v) { *(uint64_t*)a = v; }
inline uint32_t sm_ld32(smaddr_t a) { return *(const uint32_t*)a; }
inline void sm_st32(smaddr_t a, uint32_t v) { *(uint32_t*)a = v; }
inline uint32_t sm_ld16(smaddr_t a) { return *(const uint16_t*)a; }
inline void sm_st16(smaddr_t a, uint32_t v) { *(uint16_t*)a = (uint16_t)v; }
#endif

#define AB_IDX_EMPTY 0xFFFFFFFFu
#define AB_ABSENT 0x7FF8AB200AB200ABull      // coefficient slot of a key whose sum cancelled to zero (dropped, as _add2 does)

struct AccTab {
    uint32_t* idx; uint32_t mask;   // hash index (power-of-two size), cell = position in the dense list or AB_IDX_EMPTY
    int shift;                      // 64 - log2(size)
    uint64_t* key;                  // word w of entry s at key[w*cap+s]
    uint64_t* val;                  // coefficient bits, or AB_ABSENT
    int cap;
    int nd;                         // entries in use (warp-uniform)
    bool overflow;
};
AB_HD uint32_t acc_hash(const AccTab& T, const Key& k) {
    uint64_t h = k.w[0];
#if AB_KW > 1
    h ^= k.w[1] * 0xC2B2AE3D27D4EB4Full;
#endif
    h *= 0x9E3779B97F4A7C15ull;
    return (uint32_t)(h >> T.shift);
}
// Place the accumulator in the warp's shared area (attempt 0) or above `above` in the arena with room for `want` entries.
AB_HD bool acc_open(Ctx& c, AccTab& T, int attempt, int above, long long want) {
    const int W = AB_W, lane = wlane();
    wsync();
    if (attempt == 0 && c.sm) {
        c.sm_dirty = 1;
        uint32_t nidx = 64; while (nidx < AB_SM_NIDX && (long long)nidx < 2 * want) nidx <<= 1;      // small sums clear a small index
        T.idx = (uint32_t*)(c.sm + AB_SM_OFF_IDX); T.mask = nidx - 1; T.key = c.sm + AB_SM_OFF_KEY; T.val = c.sm + AB_SM_OFF_VAL;
        T.cap = (int)(nidx / 2) < AB_SM_CAP ? (int)(nidx / 2) : AB_SM_CAP;
    } else {
        if (want < 64) want = 64;
        { const long long room = ((long long)c.A.limit - above - 2) / (AB_KW + 5);      // index (<= 4 cells per entry) + keys + coefficient
          if (want > room) want = room; if (want < 1) { fail(c, ST_ARENA); return false; } }
        long long nidx = 128; while (nidx < 2 * want) nidx <<= 1;
        const long long words = nidx / 2 + (long long)(AB_KW + 1) * want;
        const long long at = (above + 1) & ~1LL;
        if (at + words > (long long)c.A.limit) { fail(c, ST_ARENA); return false; }
        T.idx = (uint32_t*)(c.A.base + at); T.mask = (uint32_t)(nidx - 1); T.key = c.A.base + at + nidx / 2; T.val = T.key + (long long)AB_KW * want; T.cap = (int)want;
        if ((int)(at + words) > c.A.peak) c.A.peak = (int)(at + words);
    }
    int lg = 0; while ((1u << lg) <= T.mask) ++lg;
    T.shift = 64 - lg;
    T.nd = 0; T.overflow = false;
    { uint32_t* const idx = T.idx; const uint32_t n = T.mask + 1u; for (uint32_t i = (uint32_t)lane; i < n; i += (uint32_t)W) idx[i] = AB_IDX_EMPTY; }
    wsync();
    return true;
}
// Collective: every lane calls it; a lane with `active` set brings a key, and the keys of the active lanes are distinct.
// Returns the key's position in the dense list, inserting it (coefficient AB_ABSENT) if it is new.
AB_HD int acc_slot(AccTab& T, bool active, const Key& k, int lane) {
    int slot = -1; uint32_t h = 0; bool isnew = false;
    if (active) {
        h = acc_hash(T, k);
        #pragma unroll 1
        for (;;) {
            const uint32_t id = T.idx[h];
            if (id == AB_IDX_EMPTY) { isnew = true; break; }
            bool eq = T.key[id] == k.w[0];
#if AB_KW > 1
            eq = eq && T.key[T.cap + id] == k.w[1];
#endif
            if (eq) { slot = (int)id; break; }
            h = (h + 1) & T.mask;
        }
    }
    const unsigned nm = wballot(isnew);
    if (nm) {
        const int cnt = popc32(nm);
        if (T.nd + cnt > T.cap) { T.overflow = true; return active ? (isnew ? 0 : slot) : -1; }   // caller checks T.overflow (uniform)
        if (isnew) {
            slot = T.nd + popc32(nm & lanemask_lt(lane));
            T.key[slot] = k.w[0];
#if AB_KW > 1
            T.key[T.cap + slot] = k.w[1];
#endif
            T.val[slot] = AB_ABSENT;
            #pragma unroll 1
            for (;;) {          // claim an empty index cell; the other new keys of this call differ from this one
                const uint32_t old = cas32(&T.idx[h], AB_IDX_EMPTY, (uint32_t)slot);
                if (old == AB_IDX_EMPTY) break;
                h = (h + 1) & T.mask;
            }
        }
        T.nd += cnt;
        wsync();
    }
    return slot;
}
// Fold one coefficient into its entry, as one step of _add2 does (src/taylor_model.cpp:505-512): returns the rounding gap
// up(u+ml) of a coincident term, 0 for a new one; a sum that cancels to zero drops the term.
AB_HD double acc_fold(AccTab& T, int slot, double tv) {
    const uint64_t cb = T.val[slot];
    if (cb == AB_ABSENT) { T.val[slot] = d2u(tv); return 0.0; }
    const double cur = u2d(cb);
    const double u = add_ru(cur, tv), ml = sub_ru(-cur, tv);
    const double nv = add_rn(cur, tv);
    T.val[slot] = (nv == 0) ? AB_ABSENT : d2u(nv);
    return add_ru(u, ml);
}
// Add the terms of model t as one stream; returns te, the upward sum of the rounding gaps in key order.
AB_HD double acc_add_model(Ctx& c, AccTab& T, const TM& t) {
    const int W = AB_W, lane = wlane(), n = t.n;
    const uint64_t* const base = c.A.base;
    double te = 0.0;
    #pragma unroll 1
    for (int j0 = 0; j0 < n; j0 += W) {
        const int j = j0 + lane; const bool act = j < n;
        Key k = key_zero(); double tv = 0.0;
        if (act) { k = bkey(base, t.off + j * AB_TW); tv = bval(base, t.off + j * AB_TW); }
        const int slot = acc_slot(T, act, k, lane);
        if (T.overflow) return te;
        double g = 0.0;
        if (act) g = acc_fold(T, slot, tv);
        const unsigned gm = wballot(g != 0.0);
        te = chain_ru_masked(te, g, gm);
    }
    wsync();
    return te;
}

// Number of monomials of degree <= D in `as` variables, saturated at 2^31-1 (also when the table is too small).
AB_HD long long monomial_count(const uint32_t* tab, int D, int as) {
    if (!tab || D + as >= AB_BINOM_N || as >= AB_BINOM_K) {
        if (!tab && D + as < 40 && as < AB_BINOM_K) { long long R = 1; for (int i = 1; i <= as && R < (1LL << 40); ++i) R = R * (D + i) / i; return R > 0x7FFFFFFFLL ? 0x7FFFFFFFLL : R; }
        return 0x7FFFFFFFLL;
    }
    return (long long)tab[(D + as) * AB_BINOM_K + as];
}
AB_HD uint32_t binom_at(const uint32_t* tab, int n, int k) { return (n < 0 || k < 0 || n >= AB_BINOM_N || k >= AB_BINOM_K) ? 0u : tab[n * AB_BINOM_K + k]; }
// Position of a monomial among all monomials in `as` variables in the graded order of include/multi_index.h:346-358.
AB_HD uint32_t key_rank(const uint32_t* tab, const Key& k, int as) {
    int R = key_deg(k);
    uint32_t r = (R > 0) ? binom_at(tab, R - 1 + as, as) : 0u;     // monomials of smaller degree
    #pragma unroll 1
    for (int t = 0; t + 1 < as; ++t) {
        const int e = key_get(k, t), m = as - t;
        r += binom_at(tab, R - e - 1 + m - 1, m - 1);              // same prefix, larger exponent here
        R -= e;
    }
    return r;
}
// Ordering of the accumulator for monomial spaces too large for the rank bitmap: bitonic sort of the entry numbers by key
// (in the hash index area, which is no longer needed). Kept out of acc_emit so that the common path stays small.
AB_HDN int acc_emit_sorted(Ctx& c, const AccTab T, int dst) {
    const int W = AB_W, lane = wlane(), nd = T.nd;
    uint64_t* const base = c.A.base;
    int P = 1; while (P < nd) P <<= 1;
    uint32_t* perm = T.idx;                       // index area holds >= 2*cap cells
    #pragma unroll 1
    for (int i = lane; i < P; i += W) perm[i] = (i < nd && T.val[i] != AB_ABSENT) ? (uint32_t)i : AB_IDX_EMPTY;
    wsync();
    #pragma unroll 1
    for (int kk = 2; kk <= P; kk <<= 1) for (int j = kk >> 1; j > 0; j >>= 1) {
        #pragma unroll 1
        for (int t = lane; t < P; t += W) {
            const int o = t ^ j;
            if (o > t) {
                const uint32_t a = perm[t], b = perm[o];
                bool b_lt_a;
                if (b == AB_IDX_EMPTY) b_lt_a = false; else if (a == AB_IDX_EMPTY) b_lt_a = true;
                else {
                    Key ka, kb; ka.w[0] = T.key[a]; kb.w[0] = T.key[b];
#if AB_KW > 1
                    ka.w[1] = T.key[T.cap + a]; kb.w[1] = T.key[T.cap + b];
#endif
                    b_lt_a = key_less(kb, ka);
                }
                const bool asc = (t & kk) == 0;
                if (b_lt_a == asc) { perm[t] = b; perm[o] = a; }
            }
        }
        wsync();
    }
    int total = 0;
    #pragma unroll 1
    for (int i0 = 0; i0 < nd; i0 += W) {
        const int i = i0 + lane; const uint32_t s = (i < nd) ? perm[i] : AB_IDX_EMPTY;
        const unsigned m = wballot(s != AB_IDX_EMPTY);
        if (s != AB_IDX_EMPTY) {
            Key k; k.w[0] = T.key[s];
#if AB_KW > 1
            k.w[1] = T.key[T.cap + s];
#endif
            bput(base, dst + (total + popc32(m & lanemask_lt(lane))) * AB_TW, k, u2d(T.val[s]));
        }
        total += popc32(m);
    }
    wsync();
    return total;
}
// Write the accumulator to the arena at word `dst` in key order, skipping dropped entries; returns the number of terms.
// `scratch` is shared (or arena) memory free for the ordering: gap-code and staging areas when the table is in shared memory.
AB_HDN int acc_emit(Ctx& c, const AccTab T, int as, int maxdeg, int dst) {
    const int W = AB_W, lane = wlane(), nd = T.nd;
    uint64_t* const base = c.A.base;
    if (nd == 0) return 0;
    if (!arena_room(c, dst, nd)) return 0;
    // rank bitmap: a bit per monomial of degree <= maxdeg; positions are prefix popcounts
    const long long R = monomial_count(c.binom, maxdeg, as);
    uint32_t* bits = 0; uint32_t* pre = 0; long long nwords = (R + 31) / 32;
    const bool in_sm = c.sm && (T.val == c.sm + AB_SM_OFF_VAL);
    if (c.binom && maxdeg + as < AB_BINOM_N && as < AB_BINOM_K && R < (1LL << 30)) {
        if (in_sm && 2 * nwords * 4 <= (long long)(AB_SM_WORDS - AB_SM_OFF_G16) * 8) { bits = (uint32_t*)(c.sm + AB_SM_OFF_G16); pre = bits + nwords; }
        else if (!in_sm) {
            // table lives in the arena: put the bitmap above it
            uint64_t* endp = T.val + T.cap;
            long long at = (endp - base); at = (at + 1) & ~1LL;
            if (nwords <= 65536 && at + nwords + 2 <= (long long)c.A.limit) { bits = (uint32_t*)(base + at); pre = bits + nwords; if ((int)(at + nwords + 2) > c.A.peak) c.A.peak = (int)(at + nwords + 2); }
        }
    }
    if (bits) {
        const uint32_t* tab = c.binom;
        #pragma unroll 1
        for (long long i = lane; i < nwords; i += W) bits[i] = 0;
        wsync();
        #pragma unroll 1
        for (int s = lane; s < nd; s += W) if (T.val[s] != AB_ABSENT) {
            Key k; k.w[0] = T.key[s];
#if AB_KW > 1
            k.w[1] = T.key[T.cap + s];
#endif
            const uint32_t r = key_rank(tab, k, as);
            or32(&bits[r >> 5], 1u << (r & 31));
        }
        wsync();
        int total = 0;
        #pragma unroll 1
        for (long long i0 = 0; i0 < nwords; i0 += W) {        // exclusive prefix of the word popcounts
            const long long i = i0 + lane; int pc = (i < nwords) ? popc32(bits[i]) : 0;
            int incl = pc;
            #pragma unroll 1
            for (int d = 1; d < W; d <<= 1) { int o = wshfl(incl, lane - d < 0 ? lane : lane - d); if (lane >= d) incl += o; }
            if (i < nwords) pre[i] = (uint32_t)(total + incl - pc);
            total += wshfl(incl, W - 1);
        }
        wsync();
        #pragma unroll 1
        for (int s = lane; s < nd; s += W) if (T.val[s] != AB_ABSENT) {
            Key k; k.w[0] = T.key[s];
#if AB_KW > 1
            k.w[1] = T.key[T.cap + s];
#endif
            const uint32_t r = key_rank(tab, k, as);
            const int pos = (int)pre[r >> 5] + popc32(bits[r >> 5] & ((1u << (r & 31)) - 1u));
            bput(base, dst + pos * AB_TW, k, u2d(T.val[s]));
        }
        wsync();
        return total;
    }
    return acc_emit_sorted(c, T, dst);
}

// ---- the dense accumulator -----------------------------------------------------------------------------------------------
// Entry r holds the monomial of rank r in the graded order (include/multi_index.h:346-358) among all monomials of degree <= D
// in `as` variables: rank = sum_{k=1..as} C(S_k - 1 + k, k), S_k = sum of the last k exponents (S_as = degree). The sums S_k are
// the bytes of (packed exponents) x 0x0101..01, the binomials come from a small table staged in shared memory, so a product
// finds its entry with one multiply and `as` table reads. An entry whose coefficient is AB_ABSENT is empty.
struct DAcc {
    smaddr_t key, val, btab;    // shared addresses: coefficient of rank r at val + 8*r, its key at key + 4*r (as <= 3) or key + 8*r
    int R;                      // ranks in use
    int as;
    uint32_t touched;           // per lane: blocks of 32 ranks this lane has written to (emission visits only those)
};
AB_HD bool dacc_fits(const Ctx& c, int as, int D) {
    if (!c.sm || !c.binom || as < 1 || as > 7 || D < 0 || D >= AB_BTAB_STRIDE) return false;
    return monomial_count(c.binom, D, as) <= (long long)(as <= 3 ? AB_SM_DENSE : AB_SM_DENSE / 2);
}
// The rank table B[k][s] = C(s-1+k, k), k = 1..7, s = 0..35, staged once per cell (ctx_stage_tables).
AB_HDN void ctx_stage_tables(Ctx& c) {
    if (!c.sm || !c.binom) return;
    const int W = AB_W, lane = wlane();
    const smaddr_t bt = sm_addr(c.sm + AB_SM_OFF_BTAB);
    wsync();
    #pragma unroll 1
    for (int q = lane; q < 7 * AB_BTAB_STRIDE; q += W) {
        const int k = q / AB_BTAB_STRIDE + 1, sk = q - (k - 1) * AB_BTAB_STRIDE;
        sm_st32(bt + 4u * (uint32_t)q, sk >= 1 ? binom_at(c.binom, sk - 1 + k, k) : 0u);
    }
    wsync();
}
// Empty every entry of the dense accumulator (at the start of a cell and after the hash accumulator has used the area).
AB_HDN void dacc_clear_all(Ctx& c) {
    const int W = AB_W, lane = wlane();
    const smaddr_t val = sm_addr(c.sm);
    wsync();
    #pragma unroll 1
    for (int r = lane; r < AB_SM_DENSE; r += W) sm_st64(val + 8u * (uint32_t)r, AB_ABSENT);
    wsync();
    c.sm_dirty = 0;
}
// All entries are empty between two uses (emission puts back what it takes), so opening costs nothing.
AB_HD void dacc_open(Ctx& c, DAcc& A, int as, int D) {
    if (c.sm_dirty) dacc_clear_all(c);
    A.val = sm_addr(c.sm); A.key = sm_addr(c.sm + AB_SM_DENSE); A.btab = sm_addr(c.sm + AB_SM_OFF_BTAB);
    A.as = as; A.R = (int)monomial_count(c.binom, D, as); A.touched = 0u;
}
AB_HD void dacc_put_key(const DAcc& A, int r, const Key& k) {
    if (A.as <= 3) sm_st32(A.key + 4u * (uint32_t)r, (uint32_t)(k.w[0] >> 32)); else sm_st64(A.key + 8u * (uint32_t)r, k.w[0]);
}
AB_HD Key dacc_get_key(const DAcc& A, int r) {
    Key k = key_zero();
    k.w[0] = (A.as <= 3) ? ((uint64_t)sm_ld32(A.key + 4u * (uint32_t)r) << 32) : sm_ld64(A.key + 8u * (uint32_t)r);
    return k;
}
AB_HD int dacc_rank(const DAcc& A, const Key& k) {
    const uint64_t S = (k.w[0] >> (8 * (7 - A.as))) * 0x0101010101010101ull;      // byte j-1 = sum of the last j exponents
    const int s1 = (int)(S & 0xFF), s2 = (int)((S >> 8) & 0xFF);
    int r = s1;                                               // C(s1, 1)
    if (A.as >= 2) r += (s2 * (s2 + 1)) >> 1;                 // C(s2 + 1, 2)
    #pragma unroll 1
    for (int j = 2; j < A.as; ++j) r += (int)sm_ld32(A.btab + 4u * (uint32_t)(j * AB_BTAB_STRIDE + (int)((S >> (8 * j)) & 0xFF)));
    return r;
}
// one step of _add2 on entry r (src/taylor_model.cpp:505-512): see acc_fold
AB_HD double dacc_fold(DAcc& A, int r, const Key& k, double tv, bool* coincident = 0) {
    const smaddr_t va = A.val + 8u * (uint32_t)r;
    A.touched |= 1u << (r >> 5);
    const uint64_t cb = sm_ld64(va);
    if (coincident) *coincident = (cb != AB_ABSENT);
    if (cb == AB_ABSENT) {
        sm_st64(va, d2u(tv));
        dacc_put_key(A, r, k);
        return 0.0;
    }
    const double cur = u2d(cb);
    const double u = add_ru(cur, tv), ml = sub_ru(-cur, tv);
    const double nv = add_rn(cur, tv);
    sm_st64(va, (nv == 0) ? AB_ABSENT : d2u(nv));
    return add_ru(u, ml);
}
// Add the terms of model t as one stream; returns te, the upward sum of the rounding gaps in key order.
AB_HDN double dacc_add_model(Ctx& c, DAcc& A, const TM& t, bool gaps) {
    const int W = AB_W, lane = wlane(), n = t.n;
    const uint64_t* const base = c.A.base;
    double te = 0.0;
    #pragma unroll 1
    for (int j0 = 0; j0 < n; j0 += W) {
        const int j = j0 + lane;
        double g = 0.0; bool co = false;
        if (j < n) { const Key k = bkey(base, t.off + j * AB_TW); g = dacc_fold(A, dacc_rank(A, k), k, bval(base, t.off + j * AB_TW), &co); }
        if (gaps) { const unsigned gm = wballot(g != 0.0); if (gm) te = chain_ru_masked(te, g, gm); }
#if defined(AB_COUNT_FLOPS) || defined(AB_MUL_HIST)
        if (gaps) { const int cc_ = popc32(wballot(co)); AB_FLOPS(5 * cc_); }
#endif
    }
    wsync();
    return te;
}
// blocks touched by any lane
AB_HD uint32_t dacc_touched(const DAcc& A, int lane) {
    uint32_t m = A.touched;
    const int W = AB_W;
    #pragma unroll 1
    for (int d = 1; d < W; d <<= 1) m |= (uint32_t)wshfl((int)m, lane ^ d);
    return m;
}
// Write the entries in use to the arena at word `dst`, in rank (= key) order, and leave the accumulator empty; returns the
// number of terms. (dst < 0: only empty it.)
AB_HDN int dacc_emit(Ctx& c, DAcc& A, int dst) {
    const int lane = wlane(), R = A.R;
    uint64_t* const base = c.A.base;
    wsync();
    uint32_t blocks = dacc_touched(A, lane);
    A.touched = 0u;
    if (dst >= 0 && !arena_room(c, dst, popc32(blocks) * 32)) dst = -1;
    int total = 0;
    #pragma unroll 1
    while (blocks) {
        const int blk = ffs32(blocks) - 1; blocks &= blocks - 1;
        #pragma unroll 1
        for (int sub = 0; sub < 32; sub += AB_W) {              // one pass on the device (32 lanes)
            const int r = blk * 32 + sub + lane;
            const smaddr_t va = A.val + 8u * (uint32_t)r;
            const uint64_t cb = (r < R) ? sm_ld64(va) : AB_ABSENT;
            const unsigned m = wballot(cb != AB_ABSENT);
            if (cb != AB_ABSENT) {
                if (dst >= 0) {
                    const int o = dst + (total + popc32(m & lanemask_lt(lane))) * AB_TW;
                    bput(base, o, dacc_get_key(A, r), u2d(cb));
                }
                sm_st64(va, AB_ABSENT);
            }
            total += popc32(m);
        }
    }
    wsync();
    return total;
}

// ---- binary operations: results are begun at the arena top ------------------------------------------------------

// r' = ((r + t[0]) + t[1]) + ... + t[B-1] with the exact semantics of B successive _add2 calls
// (src/taylor_model.cpp:487-530): every coefficient is folded in stream order; each stream has its own upward-rounded
// gap sum te[s] (visited in key order, as in the reference), and the error chain e <- up(up(e + t[s].e) + te[s]/2) is
// applied in stream order. A coefficient that cancels to exactly zero is dropped, as _add2 does, and a later stream
// re-creates the term. Result at the arena top.
#ifndef AB_MAXSTREAMS
#define AB_MAXSTREAMS 32
#endif
AB_HDN TM tm_add_many_dense(Ctx& c, const TM& r, const TM* t, int B, Acc acc, int D) {
    const int m = c.A.top;
    TM out; out.off = m; out.n = 0; out.as = r.as; out.acc = acc; out.e = r.e;
    DAcc A; dacc_open(c, A, r.as, D);
    dacc_add_model(c, A, r, false);
    double e = r.e;
    #pragma unroll 1
    for (int s = 0; s < B; ++s) { const double te = dacc_add_model(c, A, t[s], true); e = add_ru(add_ru(e, t[s].e), te / 2); }
    out.n = dacc_emit(c, A, m);
    out.e = e;
    c.A.top = m + out.n * AB_TW;
    arena_mark_peak(c);
    c.rnd = RN;
    return out;
}
AB_HDN TM tm_add_many(Ctx& c, const TM& r, const TM* t, int B, Acc acc) { AB_CALL(1, r.n); AB_CALL(9, B); AB_FLOPS(3 * B);
    const int m = c.A.top;
    TM out; out.off = m; out.n = 0; out.as = r.as; out.acc = acc; out.e = r.e;
    long long want = r.n; int maxdeg = tm_degree(c, r);
    #pragma unroll 1
    for (int s = 0; s < B; ++s) { want += t[s].n; const int d = tm_degree(c, t[s]); if (d > maxdeg) maxdeg = d; }
    if (dacc_fits(c, r.as, maxdeg)) return tm_add_many_dense(c, r, t, B, acc, maxdeg);
    double te[AB_MAXSTREAMS];
    AccTab T;
    #pragma unroll 1
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (!acc_open(c, T, attempt, m + (int)want * AB_TW, want)) return out;
        acc_add_model(c, T, r);
        #pragma unroll 1
        for (int s = 0; s < B && !T.overflow; ++s) te[s] = acc_add_model(c, T, t[s]);
        if (!T.overflow) break;
        if (attempt == 1) { fail(c, ST_ARENA); return out; }
    }
    out.n = acc_emit(c, T, r.as, maxdeg, m);
    c.A.top = m + out.n * AB_TW;
    arena_mark_peak(c);
    double e = r.e;
    #pragma unroll 1
    for (int s = 0; s < B; ++s) e = add_ru(add_ru(e, t[s].e), te[s] / 2);
    out.e = e;
    c.rnd = RN;
    return out;
}
// _add2 (src/taylor_model.cpp:487-530). Result accuracy is supplied by the caller (the reference's operator+ uses a
// fresh default).
AB_HDN TM tm_add2(Ctx& c, const TM& x, const TM& y, Acc acc) { AB_CALL(0, x.n + y.n);
    TM r = tm_add_many(c, x, &y, 1, acc);
    return r;
}
// neg (src/taylor_model.cpp:1530-1537): fresh default accuracy
AB_HDN TM tm_neg(Ctx& c, const TM& x) { AB_CALL(6, x.n);
    wsync();
    TM r = tm_begin(c, x.as, acc_default());
    if (!arena_room(c, r.off, x.n)) return r;
    uint64_t* const base = c.A.base; const int W = AB_W;
    #pragma unroll 1
    for (int i = wlane(); i < x.n; i += W) bput(base, r.off + i * AB_TW, bkey(base, x.off + i * AB_TW), -bval(base, x.off + i * AB_TW));
    wsync();
    r.n = x.n; r.e = x.e;
    c.A.top = r.off + r.n * AB_TW;
    arena_mark_peak(c);
    return r;
}
// operator+(x,y) src/taylor_model.cpp:1253-1256
AB_HD TM tm_add(Ctx& c, const TM& x, const TM& y) { return tm_add2(c, x, y, acc_default()); }
// operator-(x,y) src/taylor_model.cpp:1258-1262: r=neg(y); _acc(r,x) == _add2(s,r,x) with r keeping neg's accuracy
AB_HDN TM tm_sub(Ctx& c, const TM& x, const TM& y) {
    int m = c.A.top;
    TM ny = tm_neg(c, y);
    TM s = tm_add2(c, ny, x, acc_default());
    tm_slide(c, s, m);
    return s;
}
// x+=y (src/taylor_model.cpp:1150-1156,538-541): result keeps x's accuracy. Returns a new model at the top.
AB_HD TM tm_plus_assign(Ctx& c, const TM& x, const TM& y) { return tm_add2(c, x, y, x.acc); }

// Gap codes of a product: bit 15 = the product was discarded (its truncation error |x||y| is recomputed by the summing
// lane), else the biased exponent of the rounding gap (0 = no gap). A gap that is not 0 or a power of two cannot occur
// for coefficients in the normal range; it is reported rather than mis-encoded.
#define AB_G16_DISCARD 0x8000u
AB_HD uint32_t gap_code(double g, bool* bad) {
    const uint64_t b = d2u(g);
    if (b & 0x000FFFFFFFFFFFFFull) *bad = true;
    return (uint32_t)(b >> 52);
}
AB_HD double gap_value(uint32_t code) { return u2d((uint64_t)code << 52); }

// _mul4 (src/taylor_model.cpp:763-824) with accuracy `acc` (operator* passes the RIGHT operand's, :1265-1268).
// Result at the arena top. The x-terms are the streams: stream i is x[i]*y, its discarded products go to the stream's
// truncation error tte[i] (t.error()=tte+tte/2), its coincident products to te[i].
#ifdef AB_MUL_HIST
extern long long g_mul_hist[8][8];   // [log2 bucket of x.n][log2 bucket of y.n]
extern long long g_mul_prod[8][8];
#endif
// A product with an empty or one-term factor (constants, variables and monomials of the op tapes, nearly discarded terms of
// a composition). Adding a fixed key keeps the key order, so the product streams never meet: the result is written directly,
// compacted over the discarded products, and the error chain of the general path reduces to the truncation terms.
AB_HDN TM tm_mul_simple(Ctx& c, const TM& x, const TM& y, Acc acc) {
    uint64_t* const base = c.A.base;
    const int W = AB_W, lane = wlane();
    const int m = c.A.top;
    const int xoff = x.off, xn = x.n, yoff = y.off, yn = y.n, as = x.as;
    const double xe = x.e, ye = y.e;
    const double sweep = acc.sweep; const int maxdeg = acc.maxdeg;
    #define AB_DISCARD(k, v) (dabs(v) < sweep || key_deg(k) > maxdeg)
    TM r; r.off = m; r.n = 0; r.as = as; r.e = 0.0; r.acc = acc;
    if (xn == 0 || yn == 0) {
        // an empty factor: no streams; only the error cross terms remain (norms of the other factor)
        double xs = 0, ys = 0;
        #pragma unroll 1
        for (int i0 = 0; i0 < xn; i0 += W) { const int i = i0 + lane; xs = chain_ru(xs, i < xn ? dabs(bval(base, xoff + i * AB_TW)) : 0.0, (xn - i0 < W) ? xn - i0 : W); }
        #pragma unroll 1
        for (int j0 = 0; j0 < yn; j0 += W) { const int j = j0 + lane; ys = chain_ru(ys, j < yn ? dabs(bval(base, yoff + j * AB_TW)) : 0.0, (yn - j0 < W) ? yn - j0 : W); }
        r.e = add_ru(0.0, add_ru(add_ru(mul_ru(xs, ye), mul_ru(ys, xe)), mul_ru(xe, ye)));
        c.rnd = RN;
        return r;
    }
    if (xn == 1 || yn == 1) {
        double e = 0.0, xs = 0.0, ys = 0.0;
        int rn = 0;
        const bool one_x = (xn == 1);
        const int n = one_x ? yn : xn, soff = one_x ? yoff : xoff;
        wsync();
        const Key fa = bkey(base, one_x ? xoff : yoff); const double fv = bval(base, one_x ? xoff : yoff), afv = dabs(fv);
        const double oe = one_x ? xe : ye;          // error of the single-term factor multiplies the norm of the other
        if (!arena_room(c, m, n)) return r;
        double tte = 0.0, norm = 0.0;
        #pragma unroll 1
        for (int j0 = 0; j0 < n; j0 += W) {
            const int j = j0 + lane; const bool act = j < n;
            Key ta = key_zero(); double sv = 0.0, tv = 0.0; bool disc = false;
            if (act) {
                const int w = soff + j * AB_TW;
                sv = bval(base, w); ta = key_add(fa, bkey(base, w));
                tv = one_x ? mul_rn(fv, sv) : mul_rn(sv, fv);
                disc = AB_DISCARD(ta, tv);
            }
            const unsigned dm = wballot(disc), km = wballot(act && !disc);
            if (act && !disc) bput(base, m + (rn + popc32(km & lanemask_lt(lane))) * AB_TW, ta, tv);
            rn += popc32(km);
            if (dm) {
                if (one_x) tte = chain_ru_masked(tte, mul_ru(afv, dabs(sv)), dm);
                else { const double t1 = mul_ru(dabs(sv), afv); e = chain_ru_masked(e, add_ru(t1, t1 / 2), dm); }
            }
            if (oe != 0.0) norm = chain_ru(norm, act ? dabs(sv) : 0.0, (n - j0 < W) ? n - j0 : W);
        }
        wsync();
        if (one_x) { e = add_ru(tte, tte / 2); xs = afv; ys = norm; } else { xs = norm; ys = afv; }
        r.n = rn;
        r.e = add_ru(e, add_ru(add_ru(mul_ru(xs, ye), mul_ru(ys, xe)), mul_ru(xe, ye)));
        c.rnd = RN;
        c.A.top = m + rn * AB_TW;
        arena_mark_peak(c);
        return r;
    }
    #undef AB_DISCARD
    return r;
}
// The general product on the dense accumulator (monomial space of the result and factor y fit the shared area): stream i is
// x[i]*y, folded by all lanes at once (the keys of one stream are distinct); gap codes per (x-term, y-term), summed afterwards
// by one lane per x-term.
AB_HDN TM tm_mul_general_dense(Ctx& c, const TM& x, const TM& y, Acc acc, int D) {
    uint64_t* const base = c.A.base;
    const int W = AB_W, lane = wlane();
    const int m = c.A.top;
    const int xoff = x.off, xn = x.n, yoff = y.off, yn = y.n, as = x.as;
    const double xe = x.e, ye = y.e;
    const double sweep = acc.sweep; const int maxdeg = acc.maxdeg;
    TM r; r.off = m; r.n = 0; r.as = as; r.e = 0.0; r.acc = acc;
    DAcc A; dacc_open(c, A, as, D);
    // factor y staged as structure-of-arrays (keys | coefficients)
    const smaddr_t yk = sm_addr(c.sm + AB_SM_OFF_YK), yv = sm_addr(c.sm + AB_SM_OFF_YV), g16 = sm_addr(c.sm + AB_SM_OFF_G16);
    #pragma unroll 1
    for (int j = lane; j < yn; j += W) {
        for (int w = 0; w < AB_KW; ++w) sm_st64(yk + 8u * (uint32_t)(w * AB_SM_YCAP + j), base[yoff + j * AB_TW + w]);
        sm_st64(yv + 8u * (uint32_t)j, base[yoff + j * AB_TW + AB_KW]);
    }
    int gs = yn + ((2 - (yn & 3)) & 3); if (gs < 2) gs = 2;      // row stride = 2 mod 4: conflict-free column reads
    int Bi = AB_SM_G16 / gs; if (Bi > W) Bi = W;
    wsync();
    double re = 0.0; bool badgap = false; int kept = 0;
    #pragma unroll 1
    for (int i0 = 0; i0 < xn; i0 += Bi) {
        const int nb = (xn - i0 < Bi) ? (xn - i0) : Bi;
        Key mxa = key_zero(); double mxv = 0.0;          // lane b holds x-term i0+b
        if (lane < nb) { mxa = bkey(base, xoff + (i0 + lane) * AB_TW); mxv = bval(base, xoff + (i0 + lane) * AB_TW); }
        #pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            Key xa; for (int w = 0; w < AB_KW; ++w) xa.w[w] = wshfl(mxa.w[w], b);
            const double xv = wshfl(mxv, b);
            const smaddr_t row = g16 + 2u * (uint32_t)(b * gs);
            #pragma unroll 1
            for (int j = lane; j < yn; j += W) {
                Key ta; for (int w = 0; w < AB_KW; ++w) ta.w[w] = xa.w[w] + sm_ld64(yk + 8u * (uint32_t)(w * AB_SM_YCAP + j));
                const double tv = mul_rn(xv, u2d(sm_ld64(yv + 8u * (uint32_t)j)));
                uint32_t code = AB_G16_DISCARD;
                if (!(dabs(tv) < sweep || key_deg(ta) > maxdeg)) code = gap_code(dacc_fold(A, dacc_rank(A, ta), ta, tv), &badgap);
                sm_st16(row + 2u * (uint32_t)j, code);
            }
            wsync();          // the next x-term may meet the entries this one wrote
        }
        // per x-term sums, upward and in y order: lane b sums the row of x-term i0+b
        double tte = 0.0, te = 0.0;
        if (lane < nb) {
            const double axv = dabs(mxv);
            const smaddr_t row = g16 + 2u * (uint32_t)(lane * gs);
            #pragma unroll 1
            for (int j = 0; j < yn; ++j) {
                const uint32_t code = sm_ld16(row + 2u * (uint32_t)j);
                if (code & AB_G16_DISCARD) tte = add_ru(tte, mul_ru(axv, dabs(u2d(sm_ld64(yv + 8u * (uint32_t)j)))));
                else { if (code) te = add_ru(te, gap_value(code)); ++kept; }
            }
        }
        wsync();
        // t.error()=tte+(tte/2) (upward), then the _add2 error chain in x-term order
        #pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            const double tt = wshfl(tte, b), ee = wshfl(te, b);
            re = add_ru(add_ru(re, add_ru(tt, tt / 2)), ee / 2);
        }
    }
    if (wany(badgap)) fail(c, ST_UNSUPPORTED);
    r.n = dacc_emit(c, A, m);
#if defined(AB_COUNT_FLOPS) || defined(AB_MUL_HIST)
    {   // coincident products (5 operations each in _add2) = surviving products that did not open a new entry
        #pragma unroll 1
        for (int d = 1; d < W; d <<= 1) kept += wshfl(kept, lane ^ d);
        if (kept > r.n) AB_FLOPS(5 * (kept - r.n));
    }
#endif
    double xs = 0, ys = 0;
    #pragma unroll 1
    for (int i0 = 0; i0 < xn; i0 += W) { const int i = i0 + lane; xs = chain_ru(xs, i < xn ? dabs(bval(base, xoff + i * AB_TW)) : 0.0, (xn - i0 < W) ? xn - i0 : W); }
    #pragma unroll 1
    for (int j0 = 0; j0 < yn; j0 += W) { const int j = j0 + lane; ys = chain_ru(ys, j < yn ? dabs(bval(base, yoff + j * AB_TW)) : 0.0, (yn - j0 < W) ? yn - j0 : W); }
    r.e = add_ru(re, add_ru(add_ru(mul_ru(xs, ye), mul_ru(ys, xe)), mul_ru(xe, ye)));
    c.rnd = RN;
    c.A.top = m + r.n * AB_TW;
    arena_mark_peak(c);
    return r;
}
AB_HDN TM tm_mul_general(Ctx& c, const TM& x, const TM& y, Acc acc) {
    { int Dd = tm_degree(c, x) + tm_degree(c, y); if (Dd > acc.maxdeg) Dd = acc.maxdeg;
      if (y.n <= AB_SM_YCAP && dacc_fits(c, x.as, Dd)) return tm_mul_general_dense(c, x, y, acc, Dd); }

    uint64_t* const base = c.A.base;
    const int W = AB_W, lane = wlane();
    const int m = c.A.top;
    const int xoff = x.off, xn = x.n, yoff = y.off, yn = y.n, as = x.as;
    const double xe = x.e, ye = y.e;
    const double sweep = acc.sweep; const int maxdeg = acc.maxdeg;
    #define AB_DISCARD(k, v) (dabs(v) < sweep || key_deg(k) > maxdeg)
    TM r; r.off = m; r.n = 0; r.as = as; r.e = 0.0; r.acc = acc;
    // ---- general product ------------------------------------------------------------------------------------------
    // An upper bound of the number of result terms: distinct monomials of degree <= min(maxdeg, deg x + deg y).
    int D = tm_degree(c, x) + tm_degree(c, y); if (D > maxdeg) D = maxdeg;
    long long bound = monomial_count(c.binom, D, as);
#ifdef AB_MULG_STATS
    if (wlane() == 0) { extern long long g_mg[64][4]; int b = 0; while ((1LL << b) < bound && b < 20) ++b; g_mg[b][0]++; g_mg[b][1] += (long long)xn * yn; g_mg[40 + (as < 20 ? as : 19)][0]++; g_mg[40 + (as < 20 ? as : 19)][1] += (long long)xn * yn; if (D > g_mg[63][0]) g_mg[63][0] = D; }
#endif
    { const long long prod = (long long)xn * (long long)yn; if (prod < bound) bound = prod; }
    double re = 0.0;
    AccTab T;
    #pragma unroll 1
    for (int attempt = 0; attempt < 2; ++attempt) {
        // factor y staged in shared memory as structure-of-arrays (keys | coefficients) when it fits
        const uint64_t* yk; const uint64_t* yv; int ystr, ykw;      // word j of y: yk[j*ystr] (+ykw for the second key word), yv[j*ystr]
        if (c.sm && yn <= AB_SM_YCAP) {
            uint64_t* sk = c.sm + AB_SM_OFF_YK; uint64_t* sv = c.sm + AB_SM_OFF_YV;
            #pragma unroll 1
            for (int j = lane; j < yn; j += W) { for (int w = 0; w < AB_KW; ++w) sk[w * AB_SM_YCAP + j] = base[yoff + j * AB_TW + w]; sv[j] = base[yoff + j * AB_TW + AB_KW]; }
            yk = sk; yv = sv; ystr = 1; ykw = AB_SM_YCAP;
        } else { yk = base + yoff; yv = base + yoff + AB_KW; ystr = AB_TW; ykw = 1; }
        // gap codes: rows of `gs` codes, one row per x-term of a batch
        uint16_t* g16; int gcap;
        int gs = yn + ((2 - (yn & 3)) & 3); if (gs < 2) gs = 2;      // row stride = 2 mod 4: conflict-free column reads
        const bool g_in_sm = c.sm && gs <= AB_SM_G16 / 2;
        // without a shared area (host-side callers) or for a very long factor the gap codes go to the arena
        const long long gwords = g_in_sm ? 0 : ((long long)gs * (W < 8 ? W : 8) + 3) / 4;
        long long cap1 = ((long long)c.A.limit - m - gwords - 8) / (2 * AB_KW + 5);   // result region + table + index
        if (cap1 > bound) cap1 = bound;
        if (cap1 < 1) { fail(c, ST_ARENA); return r; }
        int tab_above = m + (int)cap1 * AB_TW;
        if (g_in_sm) { g16 = (uint16_t*)(c.sm + AB_SM_OFF_G16); gcap = AB_SM_G16; }
        else {
            const long long at = ((long long)tab_above + 1) & ~1LL;
            if (at + gwords > (long long)c.A.limit) { fail(c, ST_ARENA); return r; }
            g16 = (uint16_t*)(base + at); gcap = (int)(gwords * 4); tab_above = (int)(at + gwords);
            if (tab_above > c.A.peak) c.A.peak = tab_above;
        }
        int Bi = gcap / gs; if (Bi > W) Bi = W; if (Bi < 1) Bi = 1;
        if (!acc_open(c, T, attempt, tab_above, cap1)) return r;
        re = 0.0;
        bool badgap = false;
        #pragma unroll 1
        for (int i0 = 0; i0 < xn && !T.overflow; i0 += Bi) {
            const int nb = (xn - i0 < Bi) ? (xn - i0) : Bi;
            // lane b holds x-term i0+b
            Key mxa = key_zero(); double mxv = 0.0;
            if (lane < nb) { mxa = bkey(base, xoff + (i0 + lane) * AB_TW); mxv = bval(base, xoff + (i0 + lane) * AB_TW); }
            #pragma unroll 1
            for (int b = 0; b < nb && !T.overflow; ++b) {
                Key xa; xa.w[0] = wshfl(mxa.w[0], b);
#if AB_KW > 1
                xa.w[1] = wshfl(mxa.w[1], b);
#endif
                const double xv = wshfl(mxv, b);
                uint16_t* row = g16 + b * gs;
                #pragma unroll 1
                for (int j0 = 0; j0 < yn; j0 += W) {
                    const int j = j0 + lane; const bool act = j < yn;
                    Key ta = key_zero(); double tv = 0.0; bool disc = false;
                    if (act) {
                        Key ya; ya.w[0] = yk[j * ystr];
#if AB_KW > 1
                        ya.w[1] = yk[j * ystr + ykw];
#endif
                        ta = key_add(xa, ya); tv = mul_rn(xv, u2d(yv[j * ystr]));
                        disc = AB_DISCARD(ta, tv);
                    }
                    const int slot = acc_slot(T, act && !disc, ta, lane);
                    if (T.overflow) break;
                    if (act) {
                        uint32_t code = AB_G16_DISCARD;
                        if (!disc) { AB_FLOPS(5); code = gap_code(acc_fold(T, slot, tv), &badgap); }
                        row[j] = (uint16_t)code;
                    }
                }
                wsync();          // the next x-term may meet the entries this one wrote
            }
            if (T.overflow) break;
            // per x-term sums, upward and in y order: lane b sums the row of x-term i0+b
            double tte = 0.0, te = 0.0;
            if (lane < nb) {
                const double axv = dabs(mxv);
                const uint16_t* row = g16 + lane * gs;
                #pragma unroll 1
                for (int j = 0; j < yn; ++j) {
                    const uint32_t code = row[j];
                    if (code & AB_G16_DISCARD) tte = add_ru(tte, mul_ru(axv, dabs(u2d(yv[j * ystr]))));
                    else if (code) te = add_ru(te, gap_value(code));
                }
            }
            wsync();
            // t.error()=tte+(tte/2) (upward), then the _add2 error chain in x-term order
            #pragma unroll 1
            for (int b = 0; b < nb; ++b) {
                const double tt = wshfl(tte, b), ee = wshfl(te, b);
                const double tev = add_ru(tt, tt / 2);
                re = add_ru(add_ru(re, tev), ee / 2);
            }
        }
        if (wany(badgap)) fail(c, ST_UNSUPPORTED);
        if (!T.overflow) break;
        if (attempt == 1) { fail(c, ST_ARENA); return r; }
    }
    #undef AB_DISCARD
    r.n = acc_emit(c, T, as, D, m);
    double xs = 0, ys = 0;
    #pragma unroll 1
    for (int i0 = 0; i0 < xn; i0 += W) { const int i = i0 + lane; xs = chain_ru(xs, i < xn ? dabs(bval(base, xoff + i * AB_TW)) : 0.0, (xn - i0 < W) ? xn - i0 : W); }
    #pragma unroll 1
    for (int j0 = 0; j0 < yn; j0 += W) { const int j = j0 + lane; ys = chain_ru(ys, j < yn ? dabs(bval(base, yoff + j * AB_TW)) : 0.0, (yn - j0 < W) ? yn - j0 : W); }
    r.e = add_ru(re, add_ru(add_ru(mul_ru(xs, ye), mul_ru(ys, xe)), mul_ru(xe, ye)));
    c.rnd = RN;
    c.A.top = m + r.n * AB_TW;
    arena_mark_peak(c);
    return r;
}
AB_HD TM tm_mul(Ctx& c, const TM& x, const TM& y, Acc acc) {
#ifdef AB_MUL_HIST
    if (wlane() == 0) { int bx = 0, by = 0; while ((1 << bx) < x.n && bx < 7) ++bx; while ((1 << by) < y.n && by < 7) ++by; g_mul_hist[bx][by]++; g_mul_prod[bx][by] += (long long)x.n * y.n; }
#endif
    AB_FLOPS((double)x.n * (5.0 * y.n + 3.0) + 2.0 * (x.n + y.n) + 6.0);
    return (x.n <= 1 || y.n <= 1) ? tm_mul_simple(c, x, y, acc) : tm_mul_general(c, x, y, acc);
}
AB_HD TM tm_mul(Ctx& c, const TM& x, const TM& y) { return tm_mul(c, x, y, y.acc); }

// Order-preserving removal of the terms flagged by the caller's predicate, in place; the magnitudes of the removed terms
// are added to the error upward, in term order. Returns through x.
#define AB_FILTER_IN_PLACE(c, x, DROP_EXPR)                                                                        \
    do {                                                                                                           \
        const int W_ = AB_W, lane_ = wlane(), n_ = (x).n; int k_ = 0; double e_ = (x).e;                         \
        wsync();                                                                                                   \
        for (int i0_ = 0; i0_ < n_; i0_ += W_) {                                                                   \
            const int i_ = i0_ + lane_; const bool act_ = i_ < n_;                                                 \
            Key a = key_zero(); double v = 0.0; bool drop_ = false;                                                \
            if (act_) { a = tm_key(c, x, i_); v = tm_val(c, x, i_); drop_ = (DROP_EXPR); }                         \
            const unsigned dm_ = wballot(act_ && drop_), km_ = wballot(act_ && !drop_);                           \
            wsync();                                                                                               \
            if (act_ && !drop_) { const int p_ = k_ + popc32(km_ & lanemask_lt(lane_)); if (p_ != i_) { tm_set_key(c, x, p_, a); tm_set_val(c, x, p_, v); } } \
            k_ += popc32(km_);                                                                                     \
            if (dm_) e_ = chain_ru_masked(e_, dabs(v), dm_);                                                       \
        }                                                                                                          \
        wsync();                                                                                                   \
        (x).n = k_; (x).e = e_;                                                                                    \
    } while (0)

// Append to model r (begun at the arena top, still the top model) the terms of x for which KEEP_EXPR holds, in order, with
// key KEY_EXPR and coefficient VAL_EXPR; the expressions see the source term as `a` (Key) and `v` (double).
#define AB_FILTER_COPY(c, r, x, KEEP_EXPR, KEY_EXPR, VAL_EXPR)                                                     \
    do {                                                                                                           \
        const int W_ = AB_W, lane_ = wlane(), n_ = (x).n; int k_ = (r).n;                                        \
        wsync();                                                                                                   \
        if (arena_room(c, (r).off, k_ + n_)) {                                                                     \
            for (int i0_ = 0; i0_ < n_; i0_ += W_) {                                                               \
                const int i_ = i0_ + lane_; const bool act_ = i_ < n_;                                             \
                Key a = key_zero(); double v = 0.0; bool keep_ = false;                                            \
                if (act_) { a = tm_key(c, x, i_); v = tm_val(c, x, i_); keep_ = (KEEP_EXPR); }                     \
                const unsigned km_ = wballot(keep_);                                                               \
                if (keep_) bput((c).A.base, (r).off + (k_ + popc32(km_ & lanemask_lt(lane_))) * AB_TW, (KEY_EXPR), (VAL_EXPR)); \
                k_ += popc32(km_);                                                                                 \
            }                                                                                                      \
            wsync();                                                                                               \
            (r).n = k_; (c).A.top = (r).off + k_ * AB_TW; arena_mark_peak(c);                                      \
        }                                                                                                          \
    } while (0)

// sweep(m) src/taylor_model.cpp:972-989 -- in place; leaves the ambient mode UPWARD (sic).
AB_HDN void tm_sweep(Ctx& c, TM& x, double m) { AB_FLOPS(x.n + 2);
    AB_FILTER_IN_PLACE(c, x, !(dabs(v) >= m));
    c.rnd = RU;
}
// clean(accuracy) src/taylor_model.cpp:895-913 -- in place
AB_HDN void tm_clean(Ctx& c, TM& x, Acc acc) { AB_FLOPS(x.n + 2);
    AB_FILTER_IN_PLACE(c, x, acc_discard(acc, a, v));
    c.rnd = RN;
}
// upward sum of the coefficient magnitudes of the terms of positive degree (all terms if with_constant), starting from s
AB_HD double tm_abs_sum(Ctx& c, const TM& x, double s, bool with_constant, double* constant) {
    const int W = AB_W, lane = wlane(), n = x.n;
    double cv = 0.0;
    #pragma unroll 1
    for (int i0 = 0; i0 < n; i0 += W) {
        const int i = i0 + lane; double g = 0.0;
        if (i < n) { const double v = tm_val(c, x, i); if (!with_constant && key_deg(tm_key(c, x, i)) == 0) cv = v; else g = dabs(v); }
        s = chain_ru(s, g, (n - i0 < W) ? n - i0 : W);
        if (i0 == 0 && !with_constant) cv = wshfl(cv, 0);          // a constant term, if any, is the first term
    }
    if (constant) *constant = cv;
    return s;
}
// range() = _range1 src/taylor_model.cpp:1566-1579
AB_HDN Ivl tm_range(Ctx& c, const TM& x) { AB_FLOPS(x.n + 2); AB_CALL(7, x.n);
    double v = 0.0;
    const double t = tm_abs_sum(c, x, x.e, false, &v);
    c.rnd = RN;
    return ivl(add_rd(-t, v), add_ru(t, v));
}
// norm() src/taylor_model.cpp:2592-2600
AB_HDN double tm_norm(Ctx& c, const TM& x) { AB_FLOPS(x.n + 2);
    const double r = tm_abs_sum(c, x, x.e, true, 0);
    c.rnd = RN;
    return r;
}
// TaylorModel::operator=(Float) src/taylor_model.cpp:183-191 -> constant model begun at the top
AB_HD TM tm_const(Ctx& c, int as, double v, Acc acc) {
    TM r = tm_begin(c, as, acc);
    if (v != 0) tm_append(c, r, key_zero(), v);
    return r;
}
// TaylorModel::variable(as,j) include/taylor_model.h:232 -- default accuracy
AB_HD TM tm_variable(Ctx& c, int as, int j) {
    TM r = tm_begin(c, as, acc_default());
    tm_append(c, r, key_unit(j), 1.0);
    return r;
}
// rescale(ocd,ncd) src/taylor_model.cpp:2092-2112; x must be the top model
AB_HDN void tm_rescale(Ctx& c, TM& x, Ivl ocd, Ivl ncd) {
    double a = ocd.l, b = ocd.u, cc = ncd.l, d = ncd.u;
    int st = ST_OK;
    Ivl tmp = ivl_div(1.0, sub_ivl(b, a), &st);
    if (st != ST_OK) fail(c, st);
    Ivl alpha = ivl_mul(sub_ivl(d, cc), tmp);
    Ivl beta = ivl_mul(ivl_sub(mul_ivl(cc, b), mul_ivl(a, d)), tmp);
    tm_scal(c, x, alpha);
    tm_acc(c, x, beta);
}
// TaylorModel::scaling(as,j,d) include/taylor_model.h:234 -- default accuracy, at the top
AB_HD TM tm_scaling(Ctx& c, int as, int j, Ivl d) {
    TM r = tm_variable(c, as, j);
    tm_rescale(c, r, ivl(-1, 1), d);
    return r;
}
// unscale(tv,ivl) src/taylor_model.cpp:2554-2569 -- copy at the top
AB_HDN TM tm_unscale(Ctx& c, const TM& tv, Ivl d) {
    TM r = tm_copy(c, tv);
    Ivl cc = ivl_add(ivl(d.l / 2), ivl(d.u / 2));
    int st = ST_OK;
    Ivl s = ivl_div(2.0, ivl_sub(ivl(d.u), ivl(d.l)), &st);
    if (st != ST_OK) fail(c, st);
    tm_acc(c, r, ivl_neg(cc));
    tm_scal(c, r, s);
    return r;
}
// rescale(tv,ivl) (free function "scale") src/taylor_model.cpp:2570-2585
AB_HDN TM tm_scale_to(Ctx& c, const TM& tv, Ivl d) {
    TM r = tm_copy(c, tv);
    Ivl cc = add_ivl(d.l / 2, d.u / 2);
    Ivl s = sub_ivl(d.u / 2, d.l / 2);
    tm_scal(c, r, s);
    tm_acc(c, r, cc);
    return r;
}
// embed(x,as) src/taylor_model.cpp:2630-2633: append `extra` trailing variables. Keys are unchanged because the
// packed layout is positional (variable j always sits in the same byte).
AB_HD TM tm_embed(Ctx& c, const TM& x, int extra) { TM r = tm_copy(c, x); r.as = x.as + extra; return r; }

// antidifferentiate(k) src/taylor_model.cpp:2251-2284 -- in place
AB_HDN void tm_antidiff(Ctx& c, TM& x, int k) { AB_FLOPS(4 * x.n + 2);
    const int W = AB_W, lane = wlane(), n = x.n;
    wsync();
    double tre = 0;
    #pragma unroll 1
    for (int i0 = 0; i0 < n; i0 += W) {
        const int i = i0 + lane; double g = 0.0;
        if (i < n) {
            Key a = tm_key(c, x, i);
            double xv = tm_val(c, x, i), mxv = -xv, u, ml;
            double cc = (double)(key_get(a, k) + 1);
            if (xv >= 0) { u = div_ru(xv, cc); ml = div_ru(mxv, cc); }
            else { u = div_ru(mxv, cc); ml = div_ru(xv, cc); }
            g = add_ru(u, ml);
            key_inc(a, k, 1);
            tm_set_key(c, x, i, a);
            tm_set_val(c, x, i, div_rn(xv, cc));
        }
        tre = chain_ru(tre, g, (n - i0 < W) ? n - i0 : W);
    }
    wsync();
    x.e = add_ru(x.e, tre / 2);
    c.rnd = RN;
}
// intersection(x,y) src/taylor_model.cpp:2883-2954. Returns false (and leaves the ambient mode upward, as the
// reference's throw does) when the models are disjoint; the caller then keeps y, as unchecked_flow does.
// The union of the two key sets is built in the accumulator (x's coefficient in the entry, y's beside it); every key is
// then treated independently, and the result error is a maximum, which does not depend on the order.
// intersection on the dense accumulator: x's coefficient in the entry, y's beside it (gap-code area)
AB_HDN bool tm_intersection_dense(Ctx& c, const TM& x, const TM& y, TM& out, int D) {
    const int W = AB_W, lane = wlane();
    const uint64_t* const base = c.A.base;
    const int m = c.A.top;
    TM r = tm_begin(c, x.as, x.acc);
    const double xe = x.e, ye = y.e;
    DAcc A; dacc_open(c, A, x.as, D);
    const smaddr_t yval = sm_addr(c.sm + AB_SM_OFF_G16);
    wsync();
    #pragma unroll 1
    for (int j = lane; j < x.n; j += W) {
        const Key k = bkey(base, x.off + j * AB_TW); const int rr = dacc_rank(A, k);
        A.touched |= 1u << (rr >> 5);
        sm_st64(yval + 8u * (uint32_t)rr, d2u(0.0));
        sm_st64(A.val + 8u * (uint32_t)rr, base[x.off + j * AB_TW + AB_KW]);
        dacc_put_key(A, rr, k);
    }
    wsync();
    #pragma unroll 1
    for (int j = lane; j < y.n; j += W) {
        const Key k = bkey(base, y.off + j * AB_TW); const int rr = dacc_rank(A, k);
        A.touched |= 1u << (rr >> 5);
        if (sm_ld64(A.val + 8u * (uint32_t)rr) == AB_ABSENT) {
            sm_st64(A.val + 8u * (uint32_t)rr, d2u(0.0));
            dacc_put_key(A, rr, k);
        }
        sm_st64(yval + 8u * (uint32_t)rr, base[y.off + j * AB_TW + AB_KW]);
    }
    wsync();
    // per key: src/taylor_model.cpp:2905-2945
    double lmax = 0.0; bool bad = false;
    #pragma unroll 1
    for (uint32_t blocks = dacc_touched(A, lane); blocks; blocks &= blocks - 1)
      for (int sub = 0; sub < 32; sub += W) {
        const int rr = (ffs32(blocks) - 1) * 32 + sub + lane;
        if (rr >= A.R) continue;
        const uint64_t xb = sm_ld64(A.val + 8u * (uint32_t)rr);
        if (xb == AB_ABSENT) continue;
        const double xv = u2d(xb), yv = u2d(sm_ld64(yval + 8u * (uint32_t)rr));
        double xu = add_ru(xv, xe), yu = add_ru(yv, ye), mxl = sub_ru(xe, xv), myl = sub_ru(ye, yv);
        double u = dmin(xu, yu), ml = dmin(mxl, myl);
        const double sg = add_ru(u, ml);
        if (sg < 0) bad = true;
        lmax = dmax(lmax, sg);
        xu = add_rn(xv, xe); yu = add_rn(yv, ye); mxl = sub_rn(xe, xv); myl = sub_rn(ye, yv);
        u = dmin(xu, yu); ml = dmin(mxl, myl);
        const double rv = sub_rn(u, ml) / 2;
        sm_st64(A.val + 8u * (uint32_t)rr, (rv != 0.0) ? d2u(rv) : AB_ABSENT);
    }
    wsync();
    if (wany(bad)) { dacc_emit(c, A, -1); c.rnd = RU; c.A.top = r.off; return false; }
    double tme = 0.0;
    #pragma unroll 1
    for (int l = 0; l < W; ++l) tme = dmax(tme, wshfl(lmax, l));
    r.n = dacc_emit(c, A, m);
    c.A.top = m + r.n * AB_TW;
    arena_mark_peak(c);
    r.e = tme / 2;
    c.rnd = RN;
    out = r;
    return true;
}
AB_HDN bool tm_intersection(Ctx& c, const TM& x, const TM& y, TM& out) {
    { int Dd = tm_degree(c, x); const int d2 = tm_degree(c, y); if (d2 > Dd) Dd = d2;
      if (AB_SM_DENSE <= AB_SM_G16 / 4 && dacc_fits(c, x.as, Dd)) { AB_FLOPS(1 + 12 * (x.n + y.n)); return tm_intersection_dense(c, x, y, out, Dd); } }

    const int W = AB_W, lane = wlane();
    uint64_t* const base = c.A.base;
    const int m = c.A.top;
    TM r = tm_begin(c, x.as, x.acc);
    AB_FLOPS(1 + 12 * (x.n + y.n));
    const long long want = (long long)x.n + y.n;
    int maxdeg = tm_degree(c, x); { const int d = tm_degree(c, y); if (d > maxdeg) maxdeg = d; }
    const double xe = x.e, ye = y.e;
    AccTab T;
    bool disjoint = false; double tme = 0.0;
    #pragma unroll 1
    for (int attempt = 0; attempt < 2; ++attempt) {
        // y's coefficients beside the entries: in the gap-code area (shared) or above the table (arena)
        const long long yv_words = (attempt == 0 && c.sm) ? 0 : ((want < 64) ? 64 : want);
        const int above = m + (int)want * AB_TW + (int)yv_words + 2;
        if (!acc_open(c, T, attempt, above, want)) return false;
        uint64_t* yval;
        if (yv_words == 0) { if (T.cap > (AB_SM_WORDS - AB_SM_OFF_G16)) { T.overflow = true; continue; } yval = c.sm + AB_SM_OFF_G16; }
        else { const long long at = ((long long)m + want * AB_TW + 1) & ~1LL; if (at + yv_words > c.A.limit) { fail(c, ST_ARENA); return false; } yval = base + at; }
        #pragma unroll 1
        for (int pass = 0; pass < 2 && !T.overflow; ++pass) {
            const TM& t = pass ? y : x;
            #pragma unroll 1
            for (int j0 = 0; j0 < t.n; j0 += W) {
                const int j = j0 + lane; const bool act = j < t.n;
                Key k = key_zero(); double tv = 0.0;
                if (act) { k = bkey(base, t.off + j * AB_TW); tv = bval(base, t.off + j * AB_TW); }
                const int before = T.nd;
                const int slot = acc_slot(T, act, k, lane);
                if (T.overflow) break;
                if (act) {
                    if (pass == 0) { T.val[slot] = d2u(tv); yval[slot] = d2u(0.0); }
                    else { if (slot >= before) T.val[slot] = d2u(0.0); yval[slot] = d2u(tv); }
                }
            }
            wsync();
        }
        if (T.overflow) { if (attempt == 1) { fail(c, ST_ARENA); return false; } continue; }
        // per key: src/taylor_model.cpp:2905-2945
        double lmax = 0.0; bool bad = false;
        #pragma unroll 1
        for (int s = lane; s < T.nd; s += W) {
            const double xv = u2d(T.val[s]), yv = u2d(yval[s]);
            double xu = add_ru(xv, xe), yu = add_ru(yv, ye), mxl = sub_ru(xe, xv), myl = sub_ru(ye, yv);
            double u = dmin(xu, yu), ml = dmin(mxl, myl);
            const double sg = add_ru(u, ml);
            if (sg < 0) bad = true;
            lmax = dmax(lmax, sg);
            xu = add_rn(xv, xe); yu = add_rn(yv, ye); mxl = sub_rn(xe, xv); myl = sub_rn(ye, yv);
            u = dmin(xu, yu); ml = dmin(mxl, myl);
            const double rv = sub_rn(u, ml) / 2;
            T.val[s] = (rv != 0.0) ? d2u(rv) : AB_ABSENT;
        }
        wsync();
        disjoint = wany(bad);
        #pragma unroll 1
        for (int l = 0; l < W; ++l) tme = dmax(tme, wshfl(lmax, l));
        break;
    }
    if (disjoint) { c.rnd = RU; c.A.top = r.off; return false; }
    r.n = acc_emit(c, T, x.as, maxdeg, m);
    c.A.top = m + r.n * AB_TW;
    arena_mark_peak(c);
    r.e = tme / 2;
    c.rnd = RN;
    out = r;
    return true;
}

} } // namespace ab200
