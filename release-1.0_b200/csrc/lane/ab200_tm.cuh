// [lane build] One frontier cell per LANE: the throughput build of the evolution kernel, used only for frontiers far larger
// than the device can hold one-warp-per-cell (ab200_capi.cu picks it above AB_LANE_THRESHOLD cells). Same algorithm and
// arithmetic as the warp build one directory up, written as sequential per-lane code over lane-interleaved scratch arenas.
// ab200_tm.cuh -- TaylorModel primitives on a per-cell scratch arena.
//
// A TaylorModel is a sparse polynomial on [-1,1]^as with a uniform error bound (reference:
// include/taylor_model.h:170-190, include/expansion.h, include/multi_index.h). Here a model is a handle
// {offset,count,as,error,accuracy} into a per-cell arena of 64-bit words; terms are [key words | coefficient]
// records kept in the reference's graded order (include/multi_index.h:346-358).
//
// MultiIndex packing: AB_KW 64-bit words per key. Word 0 = degree byte (bits 56..63) followed by the exponents of
// variables 0..6, most significant first; word 1 = exponents of variables 7..14. Adding two keys is a plain
// integer add per word; the reference order is "degree ascending, then the key whose first differing exponent is
// LARGER comes first", i.e. ascending (w0 ^ 0x00FF..FF) then descending w1.
//
// Arena layout in HBM: one arena per cell, word i of cell c lives at base[(i * AB_STRIDE)] with base already offset
// by the lane; with AB_STRIDE=32 the 32 cells of a warp interleave word-by-word so that lanes walking their models
// in lock step touch one 256-byte line per access.
//
// Every accumulator keeps the reference's association order (error sums round upward term by term), so results
// are bit-identical to the CPU library; see DESIGN.md "Numerical contract".
#pragma once
#include "ab200_fp.cuh"

#ifndef AB_KW
#define AB_KW 1
#endif
#ifndef AB_STRIDE
#if defined(__CUDA_ARCH__)
#define AB_STRIDE 32
#else
#define AB_STRIDE 1
#endif
#endif

// AB_CHUNK consecutive words of one lane are kept together (AB_CHUNK=4: one 32-byte sector, 16: one 128-byte line)
// and the chunks of the 32 lanes interleave. AB_CHUNK=1 is the plain word-by-word interleave. A larger chunk keeps a
// fetched sector useful to the lane that fetched it when the lanes of a warp have drifted apart in their models.
#ifndef AB_CHUNK
#define AB_CHUNK 1
#endif
#if AB_CHUNK == 1 || AB_STRIDE == 1
#define AB_WORD_INDEX(i) ((size_t)(i) * AB_STRIDE)
#else
#define AB_WORD_INDEX(i) ((size_t)((i) / AB_CHUNK) * (AB_STRIDE * AB_CHUNK) + (size_t)((i) % AB_CHUNK))
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
    for (int j = 0; j < as; ++j) { if (j == k) continue; int e = key_get(a, j); if (e) key_inc(r, jj, e); ++jj; }
    return r;
}

// Sortable form of a key for the merge loops: unsigned comparison of (s, then t) is the graded order of key_less, and
// the all-ones value, which no monomial maps to (it would need degree 255 with all exponents 0), marks an exhausted
// stream, so that "smallest live head" is a plain running minimum.
struct SKey {
    uint64_t s;
#if AB_KW > 1
    uint64_t t;
#endif
};
AB_HD SKey skey(const Key& k) {
    SKey r; r.s = k.w[0] ^ 0x00FFFFFFFFFFFFFFull;
#if AB_KW > 1
    r.t = ~k.w[1];
#endif
    return r;
}
AB_HD SKey skey_dead() {
    SKey r; r.s = ~0ull;
#if AB_KW > 1
    r.t = ~0ull;
#endif
    return r;
}
AB_HD bool skey_is_dead(const SKey& a) { return a.s == ~0ull; }
AB_HD Key skey_key(const SKey& a) {
    Key k; k.w[0] = a.s ^ 0x00FFFFFFFFFFFFFFull;
#if AB_KW > 1
    k.w[1] = ~a.t;
#endif
    return k;
}
AB_HD bool skey_less(const SKey& a, const SKey& b) {
#if AB_KW > 1
    return a.s < b.s || (a.s == b.s && a.t < b.t);
#else
    return a.s < b.s;
#endif
}
AB_HD bool skey_eq(const SKey& a, const SKey& b) {
#if AB_KW > 1
    return a.s == b.s && a.t == b.t;
#else
    return a.s == b.s;
#endif
}

struct Acc {          // TaylorModel::Accuracy (include/taylor_model.h:178-190)
    double sweep;
    int maxdeg;
};
AB_HD Acc acc_default() { Acc a; a.sweep = 1e-18; a.maxdeg = 16; return a; }   // src/taylor_model.cpp:46-47
AB_HD bool acc_discard(const Acc& a, const Key& k, double v) { return dabs(v) < a.sweep || key_deg(k) > a.maxdeg; }

struct Arena {
    uint64_t* base;   // already offset by lane; word i at base[i*AB_STRIDE]
    int top;
    int limit;
    int peak;
};

struct Ctx {
    Arena A;
    int rnd;       // ambient rounding mode left behind by the last routine (mirrors MXCSR in the reference)
    int status;    // sticky first error
};

#ifdef AB_COUNT_ACCESS
extern long long g_aw_count;
extern long long g_sec[16];
struct SecTimer { int id; long long start; SecTimer(int i) : id(i), start(g_aw_count) { } ~SecTimer() { g_sec[id] += g_aw_count - start; } };
#define AB_SECTION(id) SecTimer sec_timer_##id(id)
AB_HD uint64_t& aw(Ctx& c, int i) { ++g_aw_count; return c.A.base[AB_WORD_INDEX(i)]; }
#else
#define AB_SECTION(id)
AB_HD uint64_t& aw(Ctx& c, int i) { return c.A.base[AB_WORD_INDEX(i)]; }
#endif
AB_HD void fail(Ctx& c, int st) { if (c.status == ST_OK) c.status = st; }
#ifdef AB_MUL_HIST
extern long long g_calls[16], g_words[16];   // development counters (tests/tools/hostdbg.cpp)
#define AB_CALL(id, w) do { g_calls[id]++; g_words[id] += (w); } while (0)
// Algorithmic FP64 operations of the REFERENCE algorithm by the per-primitive formulas of SURVEY.md 8(d), evaluated on
// the actual term counts (host development builds only; the figure bench.py quotes comes from these counters).
extern double g_flops;
#define AB_FLOPS(n) do { g_flops += (double)(n); } while (0)
#else
#define AB_CALL(id, w)
#define AB_FLOPS(n)
#endif

struct TM {
    int off;      // first word in the arena
    int n;        // number of terms
    int as;       // argument size
    double e;     // uniform error
    Acc acc;
};

AB_HD Key tm_key(Ctx& c, const TM& x, int i) { Key k; for (int j = 0; j < AB_KW; ++j) k.w[j] = aw(c, x.off + i * AB_TW + j); return k; }
AB_HD double tm_val(Ctx& c, const TM& x, int i) { return u2d(aw(c, x.off + i * AB_TW + AB_KW)); }
AB_HD void tm_set_val(Ctx& c, const TM& x, int i, double v) { aw(c, x.off + i * AB_TW + AB_KW) = d2u(v); }
AB_HD void tm_set_key(Ctx& c, const TM& x, int i, const Key& k) { for (int j = 0; j < AB_KW; ++j) aw(c, x.off + i * AB_TW + j) = k.w[j]; }
AB_HD int tm_end(const TM& x) { return x.off + x.n * AB_TW; }

// The hot loops below work on copies of the context and handle fields (arena base, limit, offsets, counts, accuracy)
// held in locals: through `Ctx&` / `const TM&` the compiler reloads them from the stack around every scratch access.
AB_HD Key bkey(const uint64_t* base, int w) { Key k; for (int j = 0; j < AB_KW; ++j) k.w[j] = base[AB_WORD_INDEX(w + j)]; return k; }
AB_HD double bval(const uint64_t* base, int w) { return u2d(base[AB_WORD_INDEX(w + AB_KW)]); }
// Streaming variants for the accumulator of a product: it is written once and read once per batch, whereas the factor y is
// re-read by every stream of every batch, so the accumulator is marked evict-first to leave the caches to y.
AB_HD Key bkey_cs(const uint64_t* base, int w) {
    Key k;
#if defined(__CUDA_ARCH__) && defined(AB_STREAMING_ACC)
    for (int j = 0; j < AB_KW; ++j) k.w[j] = __ldcs((const unsigned long long*)&base[AB_WORD_INDEX(w + j)]);
#else
    for (int j = 0; j < AB_KW; ++j) k.w[j] = base[AB_WORD_INDEX(w + j)];
#endif
    return k;
}
AB_HD double bval_cs(const uint64_t* base, int w) {
#if defined(__CUDA_ARCH__) && defined(AB_STREAMING_ACC)
    return u2d(__ldcs((const unsigned long long*)&base[AB_WORD_INDEX(w + AB_KW)]));
#else
    return u2d(base[AB_WORD_INDEX(w + AB_KW)]);
#endif
}
AB_HD void bput_cs(uint64_t* base, int w, const Key& k, double v) {
#if defined(__CUDA_ARCH__) && defined(AB_STREAMING_ACC)
    for (int j = 0; j < AB_KW; ++j) __stcs((unsigned long long*)&base[AB_WORD_INDEX(w + j)], (unsigned long long)k.w[j]);
    __stcs((unsigned long long*)&base[AB_WORD_INDEX(w + AB_KW)], (unsigned long long)d2u(v));
#else
    for (int j = 0; j < AB_KW; ++j) base[AB_WORD_INDEX(w + j)] = k.w[j];
    base[AB_WORD_INDEX(w + AB_KW)] = d2u(v);
#endif
}
AB_HD void bput(uint64_t* base, int w, const Key& k, double v) { for (int j = 0; j < AB_KW; ++j) base[AB_WORD_INDEX(w + j)] = k.w[j]; base[AB_WORD_INDEX(w + AB_KW)] = d2u(v); }

// Begin a model at the arena top; append() grows it. Only the most recently begun model may be appended to.
AB_HD TM tm_begin(Ctx& c, int as, Acc acc) { TM r; r.off = c.A.top; r.n = 0; r.as = as; r.e = 0.0; r.acc = acc; return r; }
AB_HD void tm_append(Ctx& c, TM& r, const Key& k, double v) {
    int o = r.off + r.n * AB_TW;
    if (o + AB_TW > c.A.limit) { fail(c, ST_ARENA); return; }
    for (int j = 0; j < AB_KW; ++j) aw(c, o + j) = k.w[j];
    aw(c, o + AB_KW) = d2u(v);
    r.n++;
    c.A.top = o + AB_TW;
    if (c.A.top > c.A.peak) c.A.peak = c.A.top;
}
// append() for the merge loops: the high-water mark is not maintained per term; the caller calls arena_mark_peak()
// once the model is complete (the top only grows while a model is being appended to).
AB_HD void tm_append_quick(Ctx& c, TM& r, const Key& k, double v) {
    int o = r.off + r.n * AB_TW;
    if (o + AB_TW > c.A.limit) { fail(c, ST_ARENA); return; }
    for (int j = 0; j < AB_KW; ++j) aw(c, o + j) = k.w[j];
    aw(c, o + AB_KW) = d2u(v);
    r.n++;
    c.A.top = o + AB_TW;
}
AB_HD void arena_mark_peak(Ctx& c) { if (c.A.top > c.A.peak) c.A.peak = c.A.top; }
// Move model x (anywhere at or above dst) down to word offset dst and set the arena top just past it.
AB_HD void tm_slide(Ctx& c, TM& x, int dst) { AB_CALL(2, x.off == dst ? 0 : x.n);
    if (x.off != dst) {
        const int nw = x.n * AB_TW, src = x.off;
        uint64_t* const base = c.A.base;
        for (int i = 0; i < nw; ++i) base[AB_WORD_INDEX(dst + i)] = base[AB_WORD_INDEX(src + i)];
        x.off = dst;
    }
    c.A.top = dst + x.n * AB_TW;
}
AB_HD TM tm_copy(Ctx& c, const TM& x) { AB_CALL(3, x.n);
    TM r = tm_begin(c, x.as, x.acc);
    int nw = x.n * AB_TW;
    if (r.off + nw > c.A.limit) { fail(c, ST_ARENA); return r; }
    { uint64_t* const base = c.A.base; const int dst = r.off, src = x.off;
      for (int i = 0; i < nw; ++i) base[AB_WORD_INDEX(dst + i)] = base[AB_WORD_INDEX(src + i)]; }
    r.n = x.n; r.e = x.e;
    c.A.top = r.off + nw;
    if (c.A.top > c.A.peak) c.A.peak = c.A.top;
    return r;
}
AB_HD TM tm_zero(Ctx& c, int as, Acc acc) { return tm_begin(c, as, acc); }
AB_HD int tm_degree(Ctx& c, const TM& x) { return x.n == 0 ? 0 : key_deg(tm_key(c, x, x.n - 1)); }   // taylor_model.h:361

// ---- in-place scalar operations -------------------------------------------------------------------------------

// _scal(TaylorModel&, Float) src/taylor_model.cpp:241-273
AB_HDN void tm_scal(Ctx& c, TM& r, double s) { AB_FLOPS(4 * r.n + 3); AB_CALL(4, r.n);
    if (s == 0) { r.n = 0; r.e = 0; return; }     // (arena words stay allocated; harmless)
    if (s == 1) return;
    if (s == 2 || s == 0.5) {
        for (int i = 0; i < r.n; ++i) tm_set_val(c, r, i, mul_m(c.rnd, tm_val(c, r, i), s));
        r.e = mul_m(c.rnd, r.e, dabs(s));
        return;
    }
    double te = 0, ms = -s;
    for (int i = 0; i < r.n; ++i) {
        double rv = tm_val(c, r, i);
        double u = mul_ru(rv, s), ml = mul_ru(rv, ms);
        te = add_ru(te, add_ru(u, ml));
    }
    r.e = mul_ru(r.e, dabs(s));
    r.e = add_ru(r.e, te / 2);
    for (int i = 0; i < r.n; ++i) tm_set_val(c, r, i, mul_rn(tm_val(c, r, i), s));
    c.rnd = RN;
}
// _scal(TaylorModel&, Interval) src/taylor_model.cpp:276-314
AB_HDN void tm_scal(Ctx& c, TM& r, Ivl s) { AB_FLOPS(5 * r.n + 4); AB_CALL(5, r.n);
    double te = 0, cu = s.u, mcl = -s.l;
    for (int i = 0; i < r.n; ++i) {
        double rv = tm_val(c, r, i), u, ml;
        if (rv >= 0) { u = mul_ru(rv, cu); ml = mul_ru(rv, mcl); }
        else { double mrv = -rv; u = mul_ru(mrv, mcl); ml = mul_ru(mrv, cu); }
        te = add_ru(te, add_ru(u, ml));
    }
    r.e = mul_ru(r.e, ivl_mag(s));
    r.e = add_ru(r.e, te / 2);
    double m = add_rn(s.u, s.l) / 2;
    for (int i = 0; i < r.n; ++i) tm_set_val(c, r, i, mul_rn(tm_val(c, r, i), m));
    c.rnd = RN;
}
// Insert a constant term in front of a model that is the LAST thing on the arena.
AB_HD void tm_prepend_top(Ctx& c, TM& r, double v) {
    int nw = r.n * AB_TW;
    if (r.off + nw + AB_TW > c.A.limit) { fail(c, ST_ARENA); return; }
    for (int i = nw - 1; i >= 0; --i) aw(c, r.off + AB_TW + i) = aw(c, r.off + i);
    Key z = key_zero();
    for (int j = 0; j < AB_KW; ++j) aw(c, r.off + j) = z.w[j];
    aw(c, r.off + AB_KW) = d2u(v);
    r.n++;
    c.A.top = r.off + r.n * AB_TW;
    if (c.A.top > c.A.peak) c.A.peak = c.A.top;
}
// _acc(TaylorModel&, Float) src/taylor_model.cpp:370-392. r must be the top model on the arena.
AB_HDN void tm_acc(Ctx& c, TM& r, double s) { AB_FLOPS(5);
    if (s == 0) return;
    if (r.n == 0 || key_deg(tm_key(c, r, 0)) > 0) { tm_prepend_top(c, r, s); return; }
    double rv = tm_val(c, r, 0);
    double rvu = add_ru(rv, s), mrvl = sub_ru(-rv, s);
    r.e = add_ru(r.e, add_ru(rvu, mrvl) / 2);
    tm_set_val(c, r, 0, add_rn(rv, s));
    c.rnd = RN;
}
// _acc(TaylorModel&, Interval) src/taylor_model.cpp:397-422. r must be the top model on the arena.
AB_HDN void tm_acc(Ctx& c, TM& r, Ivl s) { AB_FLOPS(6);
    if (s.l == -s.u) { r.e = add_ru(r.e, s.u); c.rnd = RN; return; }
    if (r.n == 0 || key_deg(tm_key(c, r, 0)) > 0) tm_prepend_top(c, r, 0.0);
    double rv = tm_val(c, r, 0);
    double rvu = add_ru(rv, s.u), mrvl = sub_ru(-rv, s.l);
    r.e = add_ru(r.e, add_ru(rvu, mrvl) / 2);
    double m = add_rn(s.u, s.l) / 2;
    tm_set_val(c, r, 0, add_rn(rv, m));
    c.rnd = RN;
}

// ---- binary operations: results are begun at the arena top ------------------------------------------------------

// _add2 (src/taylor_model.cpp:487-530): merge by key; rounding error of coincident terms accumulated upward in
// key order. Result accuracy is supplied by the caller (the reference's operator+ uses a fresh default).
AB_HDN TM tm_add2(Ctx& c, const TM& x, const TM& y, Acc acc) { AB_FLOPS(3); AB_CALL(0, x.n + y.n);
    TM r = tm_begin(c, x.as, acc);
    double te = 0.0;
    int i = 0, j = 0;
    bool coincident = false;
    while (i < x.n && j < y.n) {
        Key xa = tm_key(c, x, i), ya = tm_key(c, y, j);
        if (key_less(xa, ya)) { tm_append(c, r, xa, tm_val(c, x, i)); ++i; }
        else if (key_less(ya, xa)) { tm_append(c, r, ya, tm_val(c, y, j)); ++j; }
        else {
            double xv = tm_val(c, x, i), yv = tm_val(c, y, j);
            double u = add_ru(xv, yv), ml = sub_ru(-xv, yv);
            te = add_ru(te, add_ru(u, ml));
            double s = add_rn(xv, yv);
            if (s != 0) tm_append(c, r, xa, s);
            AB_FLOPS(5);
            ++i; ++j; coincident = true;
        }
    }
    while (i < x.n) { tm_append(c, r, tm_key(c, x, i), tm_val(c, x, i)); ++i; }
    while (j < y.n) { tm_append(c, r, tm_key(c, y, j), tm_val(c, y, j)); ++j; }
    (void)coincident;
    r.e = add_ru(add_ru(x.e, y.e), te / 2);
    c.rnd = RN;
    return r;
}
// r' = ((r + t[0]) + t[1]) + ... + t[B-1] with the exact semantics of B successive _add2 calls
// (src/taylor_model.cpp:487-530), computed in ONE pass over r: a (B+1)-way merge in key order in which, for every
// key, the streams are folded in index order. Each stream keeps its own upward-rounded gap sum te[s] (visited in key
// order, as in the reference), and the error chain e <- up(up(e + t[s].e) + te[s]/2) is applied in stream order.
// A coefficient that cancels to exactly zero is dropped, as _add2 does, and a later stream re-creates the term.
// This removes the O(#streams x |r|) re-copying of the accumulator that dominates the reference's _mul4/_compose2.
#ifndef AB_MAXSTREAMS
#define AB_MAXSTREAMS 8
#endif
AB_HDN TM tm_add_many(Ctx& c, const TM& r, const TM* t, int B, Acc acc) { AB_CALL(1, r.n); AB_CALL(9, B); AB_FLOPS(3 * B);
    uint64_t* const base = c.A.base;
    const int limit = c.A.limit;
    const int ooff = c.A.top;
    int on = 0;
    bool overflow = false;
    int tw[AB_MAXSTREAMS], tend[AB_MAXSTREAMS];      // word cursor and end of every stream
    double te[AB_MAXSTREAMS];
    SKey hs[AB_MAXSTREAMS];
    #pragma unroll
    for (int s = 0; s < AB_MAXSTREAMS; ++s) {
        te[s] = 0.0; hs[s] = skey_dead(); tw[s] = 0; tend[s] = 0;
        if (s < B) { tw[s] = t[s].off; tend[s] = t[s].off + t[s].n * AB_TW; if (tw[s] < tend[s]) hs[s] = skey(bkey(base, tw[s])); }
    }
    int rw = r.off; const int rend = r.off + r.n * AB_TW;
    SKey rs = rw < rend ? skey(bkey(base, rw)) : skey_dead();
    for (;;) {
        // smallest key among the heads
        SKey mk = rs;
        #pragma unroll
        for (int s = 0; s < AB_MAXSTREAMS; ++s) if (skey_less(hs[s], mk)) mk = hs[s];
        if (skey_is_dead(mk)) break;
        bool present = false;
        double cur = 0.0;
        if (skey_eq(rs, mk)) {
            cur = bval(base, rw); present = true;
            rw += AB_TW; rs = rw < rend ? skey(bkey(base, rw)) : skey_dead();
        }
        #pragma unroll
        for (int s = 0; s < AB_MAXSTREAMS; ++s) {
            if (!skey_eq(hs[s], mk)) continue;
            double tv = bval(base, tw[s]);
            if (present) {
                AB_FLOPS(5);
                double u = add_ru(cur, tv), ml = sub_ru(-cur, tv);
                te[s] = add_ru(te[s], add_ru(u, ml));
                cur = add_rn(cur, tv);
                if (cur == 0) present = false;
            } else { cur = tv; present = true; }
            tw[s] += AB_TW;
            hs[s] = tw[s] < tend[s] ? skey(bkey(base, tw[s])) : skey_dead();
        }
        if (present) { const int o = ooff + on * AB_TW; if (o + AB_TW > limit) overflow = true; else { bput(base, o, skey_key(mk), cur); ++on; } }
    }
    TM out; out.off = ooff; out.n = on; out.as = r.as; out.acc = acc;
    c.A.top = ooff + on * AB_TW;
    arena_mark_peak(c);
    if (overflow) fail(c, ST_ARENA);
    double e = r.e;
    for (int s = 0; s < B; ++s) e = add_ru(add_ru(e, t[s].e), te[s] / 2);
    out.e = e;
    c.rnd = RN;
    return out;
}
// neg (src/taylor_model.cpp:1530-1537): fresh default accuracy
AB_HDN TM tm_neg(Ctx& c, const TM& x) { AB_CALL(6, x.n);
    TM r = tm_begin(c, x.as, acc_default());
    for (int i = 0; i < x.n; ++i) tm_append(c, r, tm_key(c, x, i), -tm_val(c, x, i));
    r.e = x.e;
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

// _mul4 (src/taylor_model.cpp:763-824) with accuracy `acc` (operator* passes the RIGHT operand's, :1265-1268).
// Result at the arena top.
// Product stream b of a block is x-term (i0+b) times y; its head is the next product that is not discarded.
// Discarded products are folded into the stream's truncation error while the cursor advances, in y order, exactly
// as the inner loop of _mul4 does.
#ifndef AB_MULSTREAMS
#define AB_MULSTREAMS 8
#endif
#ifdef AB_MUL_HIST
extern long long g_mul_hist[8][8];   // [log2 bucket of x.n][log2 bucket of y.n]
extern long long g_mul_prod[8][8];
#endif
AB_HDN TM tm_mul(Ctx& c, const TM& x, const TM& y, Acc acc) {
    AB_SECTION(0);
#ifdef AB_MUL_HIST
    { int bx = 0, by = 0; while ((1 << bx) < x.n && bx < 7) ++bx; while ((1 << by) < y.n && by < 7) ++by; g_mul_hist[bx][by]++; g_mul_prod[bx][by] += (long long)x.n * y.n; }
#endif
    AB_FLOPS((double)x.n * (5.0 * y.n + 3.0) + 2.0 * (x.n + y.n) + 6.0);
    uint64_t* const base = c.A.base;
    const int limit = c.A.limit;
    const int m = c.A.top;
    const int xoff = x.off, xn = x.n, yoff = y.off, yn = y.n, as = x.as;
    const double xe = x.e, ye = y.e;
    const double sweep = acc.sweep; const int maxdeg = acc.maxdeg;
    #define AB_DISCARD(k, v) (dabs(v) < sweep || key_deg(k) > maxdeg)
    bool overflow = false;
    TM r; r.off = m; r.n = 0; r.as = as; r.e = 0.0; r.acc = acc;          // empty accumulator at m
    if (xn == 1 || yn == 1) {
        // A single-term factor (constants, variables and monomials of the op tapes: about 1400 of the 1600 products
        // of a van der Pol step). Adding a fixed key keeps the key order, so the product streams never meet: the
        // result is written directly and the error chain of the general path reduces to the truncation terms.
        double e = 0.0, xs = 0.0, ys = 0.0;
        int rn = 0;
        if (xn == 1) {
            const Key xa = bkey(base, xoff); const double xv = bval(base, xoff), axv = dabs(xv);
            double tte = 0.0;
            for (int j = 0; j < yn; ++j) {
                const int w = yoff + j * AB_TW;
                const double yv = bval(base, w); const Key ta = key_add(xa, bkey(base, w));
                const double tv = mul_rn(xv, yv);
                ys = add_ru(ys, dabs(yv));
                if (AB_DISCARD(ta, tv)) tte = add_ru(tte, mul_ru(axv, dabs(yv)));
                else { const int o = m + rn * AB_TW; if (o + AB_TW > limit) overflow = true; else { bput(base, o, ta, tv); ++rn; } }
            }
            e = add_ru(tte, tte / 2);
            xs = axv;
        } else {
            const Key ya = bkey(base, yoff); const double yv = bval(base, yoff), ayv = dabs(yv);
            for (int i = 0; i < xn; ++i) {
                const int w = xoff + i * AB_TW;
                const double xv = bval(base, w); const Key ta = key_add(bkey(base, w), ya);
                const double tv = mul_rn(xv, yv);
                xs = add_ru(xs, dabs(xv));
                if (AB_DISCARD(ta, tv)) { const double tte = mul_ru(dabs(xv), ayv); e = add_ru(e, add_ru(tte, tte / 2)); }
                else { const int o = m + rn * AB_TW; if (o + AB_TW > limit) overflow = true; else { bput(base, o, ta, tv); ++rn; } }
            }
            ys = ayv;
        }
        r.n = rn;
        r.e = add_ru(e, add_ru(add_ru(mul_ru(xs, ye), mul_ru(ys, xe)), mul_ru(xe, ye)));
        c.rnd = RN;
        c.A.top = m + rn * AB_TW;
        arena_mark_peak(c);
        if (overflow) fail(c, ST_ARENA);
        return r;
    }
    // Accumulator placement. Each batch reads the accumulator r and writes a new one; instead of building it above r and
    // sliding it back down every time, two regions [m, H) and [H, ...) are used alternately. Both hold any possible
    // accumulator: its keys are distinct monomials of degree <= min(maxdeg, deg x + deg y) in x.as variables, and there
    // are at most x.n*y.n of them. If the arena cannot hold both regions the slide per batch is kept.
    int H = -1;
    {
        int D = tm_degree(c, x) + tm_degree(c, y); if (D > maxdeg) D = maxdeg;
        long long cap = 1;
        for (int i = 1; i <= as && cap < (1LL << 24); ++i) cap = cap * (D + i) / i;
        const long long prod = (long long)xn * (long long)yn;
        if (prod < cap) cap = prod;
        const long long words = cap * AB_TW;
        if ((long long)m + 2 * words <= (long long)limit) H = m + (int)words;
    }
    const int nbatch = (xn + AB_MULSTREAMS - 1) / AB_MULSTREAMS;
    int roff = m, rn = 0, top_max = m;
    double re = 0.0;
    for (int i0 = 0; i0 < xn; i0 += AB_MULSTREAMS) {
        const int B = (xn - i0 < AB_MULSTREAMS) ? (xn - i0) : AB_MULSTREAMS;
        Key xa[AB_MULSTREAMS];
        SKey hs[AB_MULSTREAMS];
        double xv[AB_MULSTREAMS], hv[AB_MULSTREAMS], tte[AB_MULSTREAMS], te[AB_MULSTREAMS];
        int cur[AB_MULSTREAMS];        // word offset of the y term AFTER the head
        // all per-stream state is indexed with compile-time constants (fully unrolled loops) so it stays in registers
        #pragma unroll
        for (int b = 0; b < AB_MULSTREAMS; ++b) {
            tte[b] = 0.0; te[b] = 0.0; cur[b] = yoff; xv[b] = 0.0; xa[b] = key_zero(); hs[b] = skey_dead(); hv[b] = 0.0;
            if (b < B) { xa[b] = bkey(base, xoff + (i0 + b) * AB_TW); xv[b] = bval(base, xoff + (i0 + b) * AB_TW); }
        }
        const int yend = yoff + yn * AB_TW;
        // advance stream b to its next surviving product
        #define AB_ADVANCE(b) do { hs[b] = skey_dead(); const double axv_ = dabs(xv[b]); \
            while (cur[b] < yend) { const Key ta_ = key_add(xa[b], bkey(base, cur[b])); const double yv_ = bval(base, cur[b]); cur[b] += AB_TW; \
                const double tv_ = mul_rn(xv[b], yv_); \
                if (AB_DISCARD(ta_, tv_)) tte[b] = add_ru(tte[b], mul_ru(axv_, dabs(yv_))); \
                else { hs[b] = skey(ta_); hv[b] = tv_; break; } } } while (0)
        #pragma unroll
        for (int b = 0; b < AB_MULSTREAMS; ++b) if (b < B) AB_ADVANCE(b);
        // regions alternate, arranged so that the last batch lands at m and no final move is needed
        const int ooff = (H >= 0) ? ((((nbatch - 1 - i0 / AB_MULSTREAMS) & 1) == 0) ? m : H) : roff + rn * AB_TW;
        int on = 0;
        int rw = roff; const int rend = roff + rn * AB_TW;
        SKey rs = rn > 0 ? skey(bkey_cs(base, rw)) : skey_dead();
        for (;;) {
            SKey mk = rs;
            #pragma unroll
            for (int b = 0; b < AB_MULSTREAMS; ++b) if (skey_less(hs[b], mk)) mk = hs[b];
            if (skey_is_dead(mk)) break;
            bool present = false; double cv = 0.0;
            if (skey_eq(rs, mk)) { cv = bval_cs(base, rw); present = true; rw += AB_TW; rs = rw < rend ? skey(bkey_cs(base, rw)) : skey_dead(); }
            #pragma unroll
            for (int b = 0; b < AB_MULSTREAMS; ++b) {
                if (!skey_eq(hs[b], mk)) continue;
                const double tv = hv[b];
                if (present) {
                    AB_FLOPS(5);
                    double u = add_ru(cv, tv), ml = sub_ru(-cv, tv);
                    te[b] = add_ru(te[b], add_ru(u, ml));
                    cv = add_rn(cv, tv);
                    if (cv == 0) present = false;
                } else { cv = tv; present = true; }
                AB_ADVANCE(b);
            }
            if (present) { const int o = ooff + on * AB_TW; if (o + AB_TW > limit) overflow = true; else { bput_cs(base, o, skey_key(mk), cv); ++on; } }
        }
        #undef AB_ADVANCE
        if (ooff + on * AB_TW > top_max) top_max = ooff + on * AB_TW;
        AB_CALL(10, on); AB_CALL(11, B * yn);
        // t.error()=tte+(tte/2) (upward), then the _add2 error chain in x-term order
        double e = re;
        #pragma unroll
        for (int b = 0; b < AB_MULSTREAMS; ++b) if (b < B) { double tev = add_ru(tte[b], tte[b] / 2); e = add_ru(add_ru(e, tev), te[b] / 2); }
        re = e; roff = ooff; rn = on;
        if (H < 0 && roff != m) { for (int i = 0; i < rn * AB_TW; ++i) base[AB_WORD_INDEX(m + i)] = base[AB_WORD_INDEX(roff + i)]; roff = m; }
    }
    #undef AB_DISCARD
    if (roff != m) { for (int i = 0; i < rn * AB_TW; ++i) base[AB_WORD_INDEX(m + i)] = base[AB_WORD_INDEX(roff + i)]; roff = m; }
    if (top_max > c.A.peak) c.A.peak = top_max;
    if (overflow) fail(c, ST_ARENA);
    r.off = m; r.n = rn; r.e = re;
    double xs = 0, ys = 0;
    for (int i = 0; i < xn; ++i) xs = add_ru(xs, dabs(bval(base, xoff + i * AB_TW)));
    for (int j = 0; j < yn; ++j) ys = add_ru(ys, dabs(bval(base, yoff + j * AB_TW)));
    r.e = add_ru(r.e, add_ru(add_ru(mul_ru(xs, ye), mul_ru(ys, xe)), mul_ru(xe, ye)));
    c.rnd = RN;
    c.A.top = m + rn * AB_TW;
    return r;
}
AB_HD TM tm_mul(Ctx& c, const TM& x, const TM& y) { return tm_mul(c, x, y, y.acc); }

// sweep(m) src/taylor_model.cpp:972-989 -- in place; leaves the ambient mode UPWARD (sic).
AB_HDN void tm_sweep(Ctx& c, TM& x, double m) { AB_FLOPS(x.n + 2);
    int k = 0;
    for (int i = 0; i < x.n; ++i) {
        double v = tm_val(c, x, i);
        if (dabs(v) >= m) { if (k != i) { tm_set_key(c, x, k, tm_key(c, x, i)); tm_set_val(c, x, k, v); } ++k; }
        else x.e = add_ru(x.e, dabs(v));
    }
    x.n = k;
    c.rnd = RU;
}
// clean(accuracy) src/taylor_model.cpp:895-913 -- in place
AB_HDN void tm_clean(Ctx& c, TM& x, Acc acc) { AB_FLOPS(x.n + 2);
    int k = 0;
    for (int i = 0; i < x.n; ++i) {
        Key a = tm_key(c, x, i); double v = tm_val(c, x, i);
        if (acc_discard(acc, a, v)) x.e = add_ru(x.e, dabs(v));
        else { if (k != i) { tm_set_key(c, x, k, a); tm_set_val(c, x, k, v); } ++k; }
    }
    x.n = k;
    c.rnd = RN;
}
// range() = _range1 src/taylor_model.cpp:1566-1579
AB_HDN Ivl tm_range(Ctx& c, const TM& x) { AB_FLOPS(x.n + 2); AB_CALL(7, x.n);
    double t = x.e, v = 0.0;
    for (int i = 0; i < x.n; ++i) {
        if (key_deg(tm_key(c, x, i)) == 0) v = tm_val(c, x, i);
        else t = add_ru(t, dabs(tm_val(c, x, i)));
    }
    c.rnd = RN;
    return ivl(add_rd(-t, v), add_ru(t, v));
}
// norm() src/taylor_model.cpp:2592-2600
AB_HDN double tm_norm(Ctx& c, const TM& x) { AB_FLOPS(x.n + 2);
    double r = x.e;
    for (int i = 0; i < x.n; ++i) r = add_ru(r, dabs(tm_val(c, x, i)));
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
    double tre = 0;
    for (int i = 0; i < x.n; ++i) {
        Key a = tm_key(c, x, i);
        double xv = tm_val(c, x, i), mxv = -xv, u, ml;
        double cc = (double)(key_get(a, k) + 1);
        if (xv >= 0) { u = div_ru(xv, cc); ml = div_ru(mxv, cc); }
        else { u = div_ru(mxv, cc); ml = div_ru(xv, cc); }
        tre = add_ru(tre, add_ru(u, ml));
        key_inc(a, k, 1);
        tm_set_key(c, x, i, a);
        tm_set_val(c, x, i, div_rn(xv, cc));
    }
    x.e = add_ru(x.e, tre / 2);
    c.rnd = RN;
}
// intersection(x,y) src/taylor_model.cpp:2883-2954. Returns false (and leaves the ambient mode upward, as the
// reference's throw does) when the models are disjoint; the caller then keeps y, as unchecked_flow does.
AB_HDN bool tm_intersection(Ctx& c, const TM& x, const TM& y, TM& out) {
    TM r = tm_begin(c, x.as, x.acc);
    double tme = 0.0;
    int i = 0, j = 0;
    AB_FLOPS(1);
    while (i < x.n || j < y.n) { AB_FLOPS(12);
        Key a; double xv, yv;
        if (i == x.n) { a = tm_key(c, y, j); yv = tm_val(c, y, j); xv = 0.0; ++j; }
        else if (j == y.n) { a = tm_key(c, x, i); xv = tm_val(c, x, i); yv = 0.0; ++i; }
        else {
            Key xa = tm_key(c, x, i), ya = tm_key(c, y, j);
            if (key_eq(xa, ya)) { a = xa; xv = tm_val(c, x, i); yv = tm_val(c, y, j); ++i; ++j; }
            else if (key_less(xa, ya)) { a = xa; xv = tm_val(c, x, i); yv = 0.0; ++i; }
            else { a = ya; yv = tm_val(c, y, j); xv = 0.0; ++j; }
        }
        double xu = add_ru(xv, x.e), yu = add_ru(yv, y.e), mxl = sub_ru(x.e, xv), myl = sub_ru(y.e, yv);
        double u = dmin(xu, yu), ml = dmin(mxl, myl);
        double s = add_ru(u, ml);
        if (s < 0) { c.rnd = RU; c.A.top = r.off; return false; }
        tme = dmax(tme, s);
        xu = add_rn(xv, x.e); yu = add_rn(yv, y.e); mxl = sub_rn(x.e, xv); myl = sub_rn(y.e, yv);
        u = dmin(xu, yu); ml = dmin(mxl, myl);
        double rv = sub_rn(u, ml) / 2;
        if (rv != 0.0) tm_append(c, r, a, rv);
        c.rnd = RN;
    }
    r.e = tme / 2;
    out = r;
    return true;
}

} } // namespace ab200
