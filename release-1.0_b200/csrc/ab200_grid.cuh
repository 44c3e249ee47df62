// ab200_grid.cuh -- dyadic grid-paving arithmetic shared by the rasteriser kernel and the host cell sets.
//
// The reference stores a denotable set as (grid, root_cell_height, root_cell_coordinates, BDD) and grows/shrinks the
// root cell in a zig-zag pattern (reference: src/bdd_set.cpp:74-118 _merging/_splitting_coordinate and
// _compute_root_grid; :1417-1451 increase_height(Box); :155-190 _minimize_height; :455-493,1548-1574
// adjoin_outer_approximation). A BDD over path bits is equivalent to the set of depth-(height+dim*accuracy) leaf
// cells, so here a cell is identified by its absolute lattice index per coordinate at resolution 2^-accuracy of
// the base grid, and the root is tracked separately. All floating-point cell bounds are produced exactly as the
// reference does (Grid::cell at the root, src/grid.cpp:248-257, then repeated midpoint bisection,
// src/vector.cpp:246-253), because the closed-interval overlap test that decides cell membership
// (src/vector.cpp:141-147, include/numeric.h:380) is sensitive to the last bit.
#pragma once
#include "ab200_fp.cuh"

namespace ab200 { inline namespace AB_VARIANT {

#define AB_GRID_MAXN 5

struct Root {
    int height;
    int coord[AB_GRID_MAXN];
};

AB_HD int merging_coordinate(int height, int dim) { return (dim - 1) - (height % dim); }                 // bdd_set.cpp:74-76
AB_HD int splitting_coordinate(int height, int dim) { return height == 0 ? 0 : (dim - 1) - ((height - 1) % dim); } // :79-82

// Root grid of a given height (_compute_root_grid, bdd_set.cpp:85-104). The origin shifts use plain double
// subtraction (round to nearest), the doublings are exact.
AB_HD void root_grid(const double* origin, const double* lengths, int dim, int height, double* ro, double* rl) {
    for (int i = 0; i < dim; ++i) { ro[i] = origin[i]; rl[i] = lengths[i]; }
    for (int h = 0; h < height; ++h) {
        int i = merging_coordinate(h, dim);
        if ((h / dim) % 2 == 1) ro[i] = sub_rn(ro[i], rl[i]);
        rl[i] = 2.0 * rl[i];
    }
}
// Grid::cell(a) (src/grid.cpp:248-257)
AB_HD void grid_cell(const double* ro, const double* rl, const int* a, int dim, Ivl* cell) {
    for (int i = 0; i < dim; ++i)
        cell[i] = ivl(add_rn(ro[i], mul_rn(rl[i], (double)a[i])), add_rn(ro[i], mul_rn(rl[i], (double)(a[i] + 1))));
}
// Integer geometry of a root: number of merges per coordinate (log2 of the side in base cells) and the lower corner
// in base-cell units.
AB_HD void root_extent(const Root& r, int dim, int* log2side, long long* lower) {
    long long off[AB_GRID_MAXN];
    for (int i = 0; i < dim; ++i) { log2side[i] = 0; off[i] = 0; }
    for (int h = 0; h < r.height; ++h) {
        int i = merging_coordinate(h, dim);
        if ((h / dim) % 2 == 1) off[i] -= (1LL << log2side[i]);
        log2side[i]++;
    }
    for (int i = 0; i < dim; ++i) lower[i] = off[i] + (long long)r.coord[i] * (1LL << log2side[i]);
}
// One step of root growth (the body of the loops in increase_height, bdd_set.cpp:1435-1448 and :196-222).
AB_HD void root_grow(Root& r, int dim) {
    int i = merging_coordinate(r.height, dim);
    if ((r.height / dim) % 2 == 1) r.coord[i]++;
    if (r.coord[i] < 0) r.coord[i]--;
    r.coord[i] = r.coord[i] / 2;
    r.height++;
}
// One step of root shrinking towards the child containing everything (_minimize_height, bdd_set.cpp:155-190).
AB_HD void root_shrink(Root& r, int dim, bool right_child) {
    int i = splitting_coordinate(r.height, dim);
    r.height--;
    int p = (r.height / dim) % 2;
    r.coord[i] = 2 * r.coord[i] - p + (right_child ? 1 : 0);
}
// Box of a root cell.
AB_HD void root_box(const double* origin, const double* lengths, int dim, const Root& r, Ivl* cell) {
    double ro[AB_GRID_MAXN], rl[AB_GRID_MAXN];
    root_grid(origin, lengths, dim, r.height, ro, rl);
    grid_cell(ro, rl, r.coord, dim, cell);
}
// increase_height(Box) (bdd_set.cpp:1417-1451): grow until the root cell strictly covers the box.
AB_HD void root_cover(const double* origin, const double* lengths, int dim, Root& r, const Ivl* box, int max_height) {
    Ivl cell[AB_GRID_MAXN];
    for (;;) {
        root_box(origin, lengths, dim, r, cell);
        bool covers = true;
        for (int i = 0; i < dim; ++i) if (!(box[i].l > cell[i].l && box[i].u < cell[i].u)) covers = false;   // Box::covers -> inside()
        if (covers || r.height >= max_height) return;
        root_grow(r, dim);
    }
}
// Number of bisections applied to coordinate i between the root (height h) and the leaves (depth h+dim*acc).
AB_HD int splits_of(int height, int dim, int acc, int i) {
    int depth = height + dim * acc, s = splitting_coordinate(height, dim), cnt = 0;
    for (int d = 0; d < depth; ++d) { if (s == i) ++cnt; s = (s + 1) % dim; }
    return cnt;
}
// Leaf-index interval (relative to the root) of the leaves whose closed interval meets [b.l,b.u] along one
// coordinate, found by descending the bisection tree with the reference's midpoints.
AB_HD void leaf_range_1d(Ivl root, int nsplit, Ivl b, long long* lo, long long* hi) {
    double l = root.l, u = root.u; long long idx = 0;
    for (int s = 0; s < nsplit; ++s) { double m = add_rn(l, u) / 2; if (m >= b.l) { u = m; idx = 2 * idx; } else { l = m; idx = 2 * idx + 1; } }
    *lo = idx;
    l = root.l; u = root.u; idx = 0;
    for (int s = 0; s < nsplit; ++s) { double m = add_rn(l, u) / 2; if (m <= b.u) { l = m; idx = 2 * idx + 1; } else { u = m; idx = 2 * idx; } }
    *hi = idx;
}
// Bounds of leaf `idx` (nsplit bisections below `root`) along one coordinate.
AB_HD Ivl leaf_interval_1d(Ivl root, int nsplit, long long idx) {
    double l = root.l, u = root.u;
    for (int s = nsplit - 1; s >= 0; --s) { double m = add_rn(l, u) / 2; if ((idx >> s) & 1) l = m; else u = m; }
    return ivl(l, u);
}

// ---- rasterising a list of boxes into one set, as outer_approximation does for one enclosure list ---------------
// (include/hybrid_set.h:976-996: a fresh set per location, adjoin_outer_approximation per enclosure, so the root used
// for box k depends on the boxes before it.) The set's contents are summarised by the lattice bounding box of its
// cells, which is all that _minimize_height needs.
struct RunningSet {
    Root root;
    int empty;
    long long lo[AB_GRID_MAXN], hi[AB_GRID_MAXN];   // absolute fine-index bounding box of the cells so far
};
AB_HD void running_init(RunningSet& rs, int dim) {
    rs.root.height = 0; rs.empty = 1;
    for (int i = 0; i < dim; ++i) { rs.root.coord[i] = 0; rs.lo[i] = 0; rs.hi[i] = 0; }
}
// Computes the absolute fine-index ranges [lo,hi] covered by `box`; updates the running root. Returns false if the
// box is empty (nothing adjoined) or the root could not be grown far enough.
AB_HD bool raster_box(const double* origin, const double* lengths, int dim, int acc, RunningSet& rs, const Ivl* box,
                      long long* lo, long long* hi) {
    for (int i = 0; i < dim; ++i) if (box[i].l > box[i].u) return false;                    // bbox.empty()
    root_cover(origin, lengths, dim, rs.root, box, 60);
    Ivl rc[AB_GRID_MAXN];
    root_box(origin, lengths, dim, rs.root, rc);
    int l2[AB_GRID_MAXN]; long long lower[AB_GRID_MAXN];
    root_extent(rs.root, dim, l2, lower);
    for (int i = 0; i < dim; ++i) {
        if (!(box[i].l > rc[i].l && box[i].u < rc[i].u)) return false;
        int ns = l2[i] + acc;     // == splits_of(height,dim,acc,i)
        long long a, b;
        leaf_range_1d(rc[i], ns, box[i], &a, &b);
        long long base = lower[i] * (1LL << acc);
        lo[i] = base + a; hi[i] = base + b;
    }
    for (int i = 0; i < dim; ++i) {
        if (rs.empty) { rs.lo[i] = lo[i]; rs.hi[i] = hi[i]; }
        else { if (lo[i] < rs.lo[i]) rs.lo[i] = lo[i]; if (hi[i] > rs.hi[i]) rs.hi[i] = hi[i]; }
    }
    rs.empty = 0;
    // minimize_height
    while (rs.root.height > 0) {
        root_extent(rs.root, dim, l2, lower);
        int i = splitting_coordinate(rs.root.height, dim);
        long long base = lower[i] * (1LL << acc);
        long long half = (1LL << (l2[i] + acc)) / 2;
        if (rs.hi[i] < base + half) root_shrink(rs.root, dim, false);
        else if (rs.lo[i] >= base + half) root_shrink(rs.root, dim, true);
        else break;
    }
    return true;
}

// ---- packed cell keys ------------------------------------------------------------------------------------------
// key = [location | idx_{0} | idx_{1} | ... ] with per-coordinate biased absolute lattice indices; the layout is
// chosen by the host so that the restriction domain (plus a wide margin) fits.
struct KeyLayout {
    int dim;
    int acc;
    int loc_bits;
    int bits[AB_GRID_MAXN];
    long long bias[AB_GRID_MAXN];     // added to the absolute fine index to make it non-negative
};
AB_HD bool key_pack(const KeyLayout& K, int loc, const long long* idx, uint64_t* key) {
    uint64_t k = (uint64_t)loc;
    for (int i = 0; i < K.dim; ++i) {
        long long v = idx[i] + K.bias[i];
        if (v < 0 || v >= (1LL << K.bits[i])) return false;
        k = (k << K.bits[i]) | (uint64_t)v;
    }
    *key = k;
    return true;
}
AB_HD void key_unpack(const KeyLayout& K, uint64_t key, int* loc, long long* idx) {
    for (int i = K.dim - 1; i >= 0; --i) { idx[i] = (long long)(key & ((1ull << K.bits[i]) - 1)) - K.bias[i]; key >>= K.bits[i]; }
    *loc = (int)key;
}

} } // namespace ab200
