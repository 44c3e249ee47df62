// ab200_paving.cpp -- host cell sets (see ab200_paving.h).
#include "ab200_paving.h"
#include <algorithm>
#include <iterator>

namespace ab200 {

void CellSet::adjoin(const std::vector<uint64_t>& o) {
    if (o.empty()) return;
    std::vector<uint64_t> r; r.reserve(keys.size() + o.size());
    std::set_union(keys.begin(), keys.end(), o.begin(), o.end(), std::back_inserter(r));
    keys.swap(r);
}
void CellSet::remove(const CellSet& o) {
    std::vector<uint64_t> r; r.reserve(keys.size());
    std::set_difference(keys.begin(), keys.end(), o.keys.begin(), o.keys.end(), std::back_inserter(r));
    keys.swap(r);
}
void CellSet::restrict(const CellSet& o) {
    std::vector<uint64_t> r; r.reserve(std::min(keys.size(), o.keys.size()));
    std::set_intersection(keys.begin(), keys.end(), o.keys.begin(), o.keys.end(), std::back_inserter(r));
    keys.swap(r);
}

Root CellSet::minimal_root() const {
    const int dim = grid.dim;
    long long lo[AB_GRID_MAXN], hi[AB_GRID_MAXN], idx[AB_GRID_MAXN];
    bool first = true;
    for (size_t k = 0; k < keys.size(); ++k) {
        int l; key_unpack(K, keys[k], &l, idx);
        for (int i = 0; i < dim; ++i) { if (first || idx[i] < lo[i]) lo[i] = idx[i]; if (first || idx[i] > hi[i]) hi[i] = idx[i]; }
        first = false;
    }
    // start from the height-0 root containing the first corner and grow until both corners are inside
    Root r; r.height = 0;
    for (int i = 0; i < dim; ++i) { long long c = lo[i] >> K.acc; r.coord[i] = (int)c; }   // floor division by 2^acc
    for (;;) {
        int l2[AB_GRID_MAXN]; long long lower[AB_GRID_MAXN];
        root_extent(r, dim, l2, lower);
        bool ok = true;
        for (int i = 0; i < dim; ++i) {
            long long base = lower[i] * (1LL << K.acc), side = 1LL << (l2[i] + K.acc);
            if (lo[i] < base || hi[i] >= base + side) ok = false;
        }
        if (ok) break;
        root_grow(r, dim);
    }
    return r;
}

namespace {
typedef unsigned __int128 Path;
struct PathCell { Path p; };
}

// Path (sequence of left/right choices from the root to the leaf) of a cell given its root-relative leaf indices.
static Path make_path(const Root& root, int dim, int acc, const long long* rel, const int* nsplit) {
    int used[AB_GRID_MAXN];
    for (int i = 0; i < dim; ++i) used[i] = 0;
    int depth = root.height + dim * acc;
    int s = splitting_coordinate(root.height, dim);
    Path p = 0;
    for (int d = 0; d < depth; ++d) {
        int bit = (int)((rel[s] >> (nsplit[s] - 1 - used[s])) & 1);
        used[s]++;
        p = (p << 1) | (Path)bit;
        s = (s + 1) % dim;
    }
    return p;
}

static void enum_rec(const std::vector<Path>& P, size_t lo, size_t hi, int level, int depth, int split, int dim,
                     CellBox box, int mince_levels, std::vector<CellBox>& out) {
    if (hi == lo) return;
    size_t full = (depth - level >= 63) ? (size_t)-1 : ((size_t)1 << (depth - level));
    if (hi - lo == full && level >= mince_levels) { out.push_back(box); return; }
    if (level == depth) { out.push_back(box); return; }
    // first index whose bit at this level is 1
    int shift = depth - 1 - level;
    size_t a = lo, b = hi;
    while (a < b) { size_t m = (a + b) / 2; if ((P[m] >> shift) & 1) b = m; else a = m + 1; }
    double l = box.b[2 * split], u = box.b[2 * split + 1];
    double c = (l + u) / 2.0;                      // tail.cell[i].set_upper((lower+upper)/2.0)  bdd_set.cpp:833,877
    CellBox left = box, right = box;
    left.b[2 * split + 1] = c; right.b[2 * split] = c;
    enum_rec(P, lo, a, level + 1, depth, (split + 1) % dim, dim, left, mince_levels, out);
    enum_rec(P, a, hi, level + 1, depth, (split + 1) % dim, dim, right, mince_levels, out);
}

void CellSet::enumerate(std::vector<CellBox>& out, int mince_levels) const {
    if (keys.empty()) return;
    const int dim = grid.dim;
    Root root = minimal_root();
    int l2[AB_GRID_MAXN], ns[AB_GRID_MAXN]; long long lower[AB_GRID_MAXN];
    root_extent(root, dim, l2, lower);
    for (int i = 0; i < dim; ++i) ns[i] = l2[i] + K.acc;
    std::vector<Path> P(keys.size());
    long long idx[AB_GRID_MAXN], rel[AB_GRID_MAXN];
    for (size_t k = 0; k < keys.size(); ++k) {
        int l; key_unpack(K, keys[k], &l, idx);
        for (int i = 0; i < dim; ++i) rel[i] = idx[i] - lower[i] * (1LL << K.acc);
        P[k] = make_path(root, dim, K.acc, rel, ns);
    }
    std::sort(P.begin(), P.end());
    Ivl rc[AB_GRID_MAXN];
    root_box(grid.origin, grid.lengths, dim, root, rc);
    CellBox rb;
    for (int i = 0; i < dim; ++i) { rb.b[2 * i] = rc[i].l; rb.b[2 * i + 1] = rc[i].u; }
    int depth = root.height + dim * K.acc;
    enum_rec(P, 0, P.size(), 0, depth, splitting_coordinate(root.height, dim), dim, rb, mince_levels, out);
}

bool CellSet::adjoin_outer_approximation(const double* box) {
    const int dim = grid.dim;
    // The reference grows the CURRENT root of the set to cover the box, rasterises, then minimises; the leaf indices
    // produced depend on the root only through the floating-point cell bounds.
    RunningSet rs; running_init(rs, dim);
    if (!keys.empty()) {
        rs.root = minimal_root(); rs.empty = 0;
        long long idx[AB_GRID_MAXN]; bool first = true;
        for (size_t k = 0; k < keys.size(); ++k) {
            int l; key_unpack(K, keys[k], &l, idx);
            for (int i = 0; i < dim; ++i) { if (first || idx[i] < rs.lo[i]) rs.lo[i] = idx[i]; if (first || idx[i] > rs.hi[i]) rs.hi[i] = idx[i]; }
            first = false;
        }
    }
    Ivl b[AB_GRID_MAXN];
    for (int i = 0; i < dim; ++i) b[i] = ivl(box[2 * i], box[2 * i + 1]);
    long long lo[AB_GRID_MAXN], hi[AB_GRID_MAXN], idx[AB_GRID_MAXN];
    if (!raster_box(grid.origin, grid.lengths, dim, K.acc, rs, b, lo, hi)) return true;
    std::vector<uint64_t> add;
    bool ok = true;
    for (int i = 0; i < dim; ++i) idx[i] = lo[i];
    for (;;) {
        uint64_t key;
        if (key_pack(K, loc, idx, &key)) add.push_back(key); else ok = false;
        int d = dim - 1;
        while (d >= 0) { if (++idx[d] <= hi[d]) break; idx[d] = lo[d]; --d; }
        if (d < 0) break;
    }
    std::sort(add.begin(), add.end());
    adjoin(add);
    return ok;
}

namespace {
struct PathKey { Path p; uint64_t key; bool operator<(const PathKey& o) const { return p < o.p; } };
}
static void restrict_rec(const std::vector<PathKey>& P, size_t lo, size_t hi, int level, int depth, int split, int dim,
                         const CellBox& box, const CellSet::BoxTest& check, std::vector<uint64_t>& out, int stop = -1) {
    if (hi == lo) return;                                   // enabled_cells == bddfalse
    int prop = check(box.b);
    if (prop == 0) return;
    if (prop == 1 || level == (stop >= 0 ? stop : depth)) { for (size_t k = lo; k < hi; ++k) out.push_back(P[k].key); return; }
    int shift = depth - 1 - level;
    size_t a = lo, b = hi;
    while (a < b) { size_t m = (a + b) / 2; if ((P[m].p >> shift) & 1) b = m; else a = m + 1; }
    double c = (box.b[2 * split] + box.b[2 * split + 1]) / 2.0;          // Box::split -> midpoint (src/vector.cpp:246-253)
    CellBox left = box, right = box;
    left.b[2 * split + 1] = c; right.b[2 * split] = c;
    restrict_rec(P, lo, a, level + 1, depth, (split + 1) % dim, dim, left, check, out, stop);
    restrict_rec(P, a, hi, level + 1, depth, (split + 1) % dim, dim, right, check, out, stop);
}

void CellSet::outer_restrict(const BoxTest& check, bool to_bdd_depth) {
    if (keys.empty()) return;
    const int dim = grid.dim;
    Root root = minimal_root();
    int l2[AB_GRID_MAXN], ns[AB_GRID_MAXN]; long long lower[AB_GRID_MAXN];
    root_extent(root, dim, l2, lower);
    for (int i = 0; i < dim; ++i) ns[i] = l2[i] + K.acc;
    std::vector<PathKey> P(keys.size());
    long long idx[AB_GRID_MAXN], rel[AB_GRID_MAXN];
    for (size_t k = 0; k < keys.size(); ++k) {
        int l; key_unpack(K, keys[k], &l, idx);
        for (int i = 0; i < dim; ++i) rel[i] = idx[i] - lower[i] * (1LL << K.acc);
        P[k].p = make_path(root, dim, K.acc, rel, ns); P[k].key = keys[k];
    }
    std::sort(P.begin(), P.end());
    Ivl rc[AB_GRID_MAXN];
    root_box(grid.origin, grid.lengths, dim, root, rc);
    CellBox rb;
    for (int i = 0; i < dim; ++i) { rb.b[2 * i] = rc[i].l; rb.b[2 * i + 1] = rc[i].u; }
    std::vector<uint64_t> out;
    const int full = root.height + dim * K.acc;
    int stop = -1;
    if (to_bdd_depth) {
        // the last variable in the support of the BDD: the deepest path bit the set is not invariant under
        stop = root.height;                                   // (depth() is never negative: at least down to the unit cells)
        for (int sh = 0; sh < full - root.height; ++sh) {
            bool invariant = true;
            for (size_t k = 0; k < P.size() && invariant; ++k) {
                PathKey q; q.p = P[k].p ^ ((Path)1 << sh); q.key = 0;
                std::vector<PathKey>::const_iterator it = std::lower_bound(P.begin(), P.end(), q);
                if (it == P.end() || it->p != q.p) invariant = false;
            }
            if (!invariant) { stop = full - sh; break; }
        }
    }
    restrict_rec(P, 0, P.size(), 0, full, splitting_coordinate(root.height, dim), dim, rb, check, out, stop);
    std::sort(out.begin(), out.end());
    keys.swap(out);
}

void CellSet::outer_restrict_batched(const BatchTest& check) {
    if (keys.empty()) return;
    const int dim = grid.dim;
    Root root = minimal_root();
    int l2[AB_GRID_MAXN], ns[AB_GRID_MAXN]; long long lower[AB_GRID_MAXN];
    root_extent(root, dim, l2, lower);
    for (int i = 0; i < dim; ++i) ns[i] = l2[i] + K.acc;
    std::vector<PathKey> P(keys.size());
    long long idx[AB_GRID_MAXN], rel[AB_GRID_MAXN];
    for (size_t k = 0; k < keys.size(); ++k) {
        int l; key_unpack(K, keys[k], &l, idx);
        for (int i = 0; i < dim; ++i) rel[i] = idx[i] - lower[i] * (1LL << K.acc);
        P[k].p = make_path(root, dim, K.acc, rel, ns); P[k].key = keys[k];
    }
    std::sort(P.begin(), P.end());
    Ivl rc[AB_GRID_MAXN];
    root_box(grid.origin, grid.lengths, dim, root, rc);
    struct Node { size_t lo, hi; int split; CellBox box; };
    std::vector<Node> level(1), next;
    level[0].lo = 0; level[0].hi = P.size(); level[0].split = splitting_coordinate(root.height, dim);
    for (int i = 0; i < dim; ++i) { level[0].box.b[2 * i] = rc[i].l; level[0].box.b[2 * i + 1] = rc[i].u; }
    const int depth = root.height + dim * K.acc;
    std::vector<uint64_t> out;
    std::vector<CellBox> boxes; std::vector<int> tri;
    for (int lv = 0; lv <= depth && !level.empty(); ++lv) {
        boxes.resize(level.size());
        for (size_t q = 0; q < level.size(); ++q) boxes[q] = level[q].box;
        check(boxes, tri);
        next.clear();
        for (size_t q = 0; q < level.size(); ++q) {
            const Node& nd = level[q];
            const int prop = tri[q];
            if (prop == 0) continue;
            if (prop == 1 || lv == depth) { for (size_t k = nd.lo; k < nd.hi; ++k) out.push_back(P[k].key); continue; }
            const int shift = depth - 1 - lv;
            size_t a = nd.lo, b = nd.hi;
            while (a < b) { size_t m = (a + b) / 2; if ((P[m].p >> shift) & 1) b = m; else a = m + 1; }
            const double c = (nd.box.b[2 * nd.split] + nd.box.b[2 * nd.split + 1]) / 2.0;
            Node L = nd, R = nd;
            L.hi = a; R.lo = a;
            L.box.b[2 * nd.split + 1] = c; R.box.b[2 * nd.split] = c;
            L.split = R.split = (nd.split + 1) % dim;
            if (L.hi > L.lo) next.push_back(L);                     // enabled_cells == bddfalse: nothing below
            if (R.hi > R.lo) next.push_back(R);
        }
        level.swap(next);
    }
    std::sort(out.begin(), out.end());
    keys.swap(out);
}

namespace {
struct AdjoinCtx {
    const CellSet* cs; const CellSet::BoxTest* disjoint; int dim, depth; long long base[AB_GRID_MAXN];
    std::vector<uint64_t> add; bool ok;
};
}
static void adjoin_rec(AdjoinCtx& A, int level, int split, const CellBox& box, long long* rel) {
    if ((*A.disjoint)(box.b) == 1) return;
    if (level == A.depth) {
        long long idx[AB_GRID_MAXN]; uint64_t key;
        for (int i = 0; i < A.dim; ++i) idx[i] = A.base[i] + rel[i];
        if (key_pack(A.cs->K, A.cs->loc, idx, &key)) A.add.push_back(key); else A.ok = false;
        return;
    }
    double c = (box.b[2 * split] + box.b[2 * split + 1]) / 2.0;
    CellBox left = box, right = box;
    left.b[2 * split + 1] = c; right.b[2 * split] = c;
    long long save = rel[split];
    rel[split] = 2 * save;     adjoin_rec(A, level + 1, (split + 1) % A.dim, left, rel);
    rel[split] = 2 * save + 1; adjoin_rec(A, level + 1, (split + 1) % A.dim, right, rel);
    rel[split] = save;
}

bool CellSet::adjoin_outer_approximation(const double* bbox, const BoxTest& disjoint) {
    const int dim = grid.dim;
    Ivl b[AB_GRID_MAXN];
    for (int i = 0; i < dim; ++i) { b[i] = ivl(bbox[2 * i], bbox[2 * i + 1]); if (b[i].l > b[i].u) return true; }   // bbox.empty()
    Root root; root.height = 0; for (int i = 0; i < dim; ++i) root.coord[i] = 0;
    if (!keys.empty()) root = minimal_root();
    root_cover(grid.origin, grid.lengths, dim, root, b, 60);
    Ivl rc[AB_GRID_MAXN];
    root_box(grid.origin, grid.lengths, dim, root, rc);
    int l2[AB_GRID_MAXN]; long long lower[AB_GRID_MAXN];
    root_extent(root, dim, l2, lower);
    AdjoinCtx A; A.cs = this; A.disjoint = &disjoint; A.dim = dim; A.depth = root.height + dim * K.acc; A.ok = true;
    for (int i = 0; i < dim; ++i) A.base[i] = lower[i] * (1LL << K.acc);
    CellBox rb; long long rel[AB_GRID_MAXN];
    for (int i = 0; i < dim; ++i) { rb.b[2 * i] = rc[i].l; rb.b[2 * i + 1] = rc[i].u; rel[i] = 0; }
    adjoin_rec(A, 0, splitting_coordinate(root.height, dim), rb, rel);
    std::sort(A.add.begin(), A.add.end());
    adjoin(A.add);
    return A.ok;
}

namespace {
struct TreeQuery {
    const std::vector<Path>* P; int depth, dim;
};
}
static void sorted_paths(const CellSet& cs, Root& root, std::vector<Path>& P, CellBox& rb) {
    const int dim = cs.grid.dim;
    root = cs.minimal_root();
    int l2[AB_GRID_MAXN], ns[AB_GRID_MAXN]; long long lower[AB_GRID_MAXN];
    root_extent(root, dim, l2, lower);
    for (int i = 0; i < dim; ++i) ns[i] = l2[i] + cs.K.acc;
    P.resize(cs.keys.size());
    long long idx[AB_GRID_MAXN], rel[AB_GRID_MAXN];
    for (size_t k = 0; k < cs.keys.size(); ++k) {
        int l; key_unpack(cs.K, cs.keys[k], &l, idx);
        for (int i = 0; i < dim; ++i) rel[i] = idx[i] - lower[i] * (1LL << cs.K.acc);
        P[k] = make_path(root, dim, cs.K.acc, rel, ns);
    }
    std::sort(P.begin(), P.end());
    Ivl rc[AB_GRID_MAXN];
    root_box(cs.grid.origin, cs.grid.lengths, dim, root, rc);
    for (int i = 0; i < dim; ++i) { rb.b[2 * i] = rc[i].l; rb.b[2 * i + 1] = rc[i].u; }
}
static bool superset_rec(const TreeQuery& Q, size_t lo, size_t hi, int level, int split, const CellBox& cell, CellBox box) {
    if (hi == lo) return false;                                           // bddfalse
    size_t full = (Q.depth - level >= 63) ? (size_t)-1 : ((size_t)1 << (Q.depth - level));
    if (hi - lo == full) return true;                                     // bddtrue
    const std::vector<Path>& P = *Q.P;
    int shift = Q.depth - 1 - level;
    size_t a = lo, b = hi;
    while (a < b) { size_t m = (a + b) / 2; if ((P[m] >> shift) & 1) b = m; else a = m + 1; }
    double c = (cell.b[2 * split] + cell.b[2 * split + 1]) / 2.0;
    CellBox left = cell, right = cell;
    left.b[2 * split + 1] = c; right.b[2 * split] = c;
    CellBox lb = box, rb = box;                                           // intersection of the box with the two subcells
    lb.b[2 * split] = std::max(box.b[2 * split], left.b[2 * split]); lb.b[2 * split + 1] = std::min(box.b[2 * split + 1], left.b[2 * split + 1]);
    rb.b[2 * split] = std::max(box.b[2 * split], right.b[2 * split]); rb.b[2 * split + 1] = std::min(box.b[2 * split + 1], right.b[2 * split + 1]);
    const int next = (split + 1) % Q.dim;
    if (!(lb.b[2 * split] > lb.b[2 * split + 1])) { if (!superset_rec(Q, lo, a, level + 1, next, left, lb)) return false; }
    if (!(rb.b[2 * split] > rb.b[2 * split + 1])) return superset_rec(Q, a, hi, level + 1, next, right, rb);
    return true;
}
CellTree::CellTree(const CellSet& cs) : depth(0), dim(cs.grid.dim), first_split(0) {
    if (cs.keys.empty()) return;
    Root root;
    sorted_paths(cs, root, paths, root_cell);
    depth = root.height + dim * cs.K.acc;
    first_split = splitting_coordinate(root.height, dim);
}
bool CellTree::superset(const double* box) const {
    for (int i = 0; i < dim; ++i) if (box[2 * i] > box[2 * i + 1]) return true;                  // empty box
    if (paths.empty()) return false;       // an empty set: outside its unit root the subset test fails, inside it meets bddfalse
    for (int i = 0; i < dim; ++i) if (!(box[2 * i] >= root_cell.b[2 * i] && box[2 * i + 1] <= root_cell.b[2 * i + 1])) return false;   // box.subset(root_cell)
    TreeQuery Q; Q.P = &paths; Q.depth = depth; Q.dim = dim;
    CellBox bx; for (int i = 0; i < 2 * dim; ++i) bx.b[i] = box[i];
    return superset_rec(Q, 0, paths.size(), 0, first_split, root_cell, bx);
}
static bool disjoint_rec(const TreeQuery& Q, size_t lo, size_t hi, int level, int split, const CellBox& cell, const double* box) {
    if (hi == lo) return true;                                            // bddfalse
    bool dis = false;
    for (int i = 0; i < Q.dim; ++i) if (box[2 * i] > cell.b[2 * i + 1] || box[2 * i + 1] < cell.b[2 * i]) dis = true;   // Box::disjoint (closed)
    if (dis) return true;
    size_t full = (Q.depth - level >= 63) ? (size_t)-1 : ((size_t)1 << (Q.depth - level));
    if (hi - lo == full) return false;                                    // bddtrue: the (false) result of the test
    const std::vector<Path>& P = *Q.P;
    int shift = Q.depth - 1 - level;
    size_t a = lo, b = hi;
    while (a < b) { size_t m = (a + b) / 2; if ((P[m] >> shift) & 1) b = m; else a = m + 1; }
    double c = (cell.b[2 * split] + cell.b[2 * split + 1]) / 2.0;
    CellBox left = cell, right = cell;
    left.b[2 * split + 1] = c; right.b[2 * split] = c;
    const int next = (split + 1) % Q.dim;
    if (!disjoint_rec(Q, lo, a, level + 1, next, left, box)) return false;
    return disjoint_rec(Q, a, hi, level + 1, next, right, box);
}
bool CellTree::disjoint(const double* box) const {
    for (int i = 0; i < dim; ++i) if (box[2 * i] > box[2 * i + 1]) return true;
    if (paths.empty()) return true;
    TreeQuery Q; Q.P = &paths; Q.depth = depth; Q.dim = dim;
    return disjoint_rec(Q, 0, paths.size(), 0, first_split, root_cell, box);
}

bool CellSet::bounding_box(double* box) const {
    std::vector<CellBox> cells; enumerate(cells);
    if (cells.empty()) return false;
    const int dim = grid.dim;
    for (int i = 0; i < dim; ++i) { box[2 * i] = cells[0].b[2 * i]; box[2 * i + 1] = cells[0].b[2 * i + 1]; }
    for (size_t k = 1; k < cells.size(); ++k)
        for (int i = 0; i < dim; ++i) { box[2 * i] = std::min(box[2 * i], cells[k].b[2 * i]); box[2 * i + 1] = std::max(box[2 * i + 1], cells[k].b[2 * i + 1]); }
    return true;
}

} // namespace ab200
