// ab200_paving.h -- host-side cell sets on a dyadic grid (the counterpart of the reference's BDDTreeSet).
//
// A CellSet is a sorted, duplicate-free array of packed lattice keys (ab200_grid.cuh: KeyLayout) for ONE location,
// i.e. exactly what the GPU sort emits, plus the grid. Union / difference / intersection are linear merges over the
// sorted arrays (reference: BDDTreeSet::adjoin/remove/restrict = bdd_or / bddop_diff / bdd_and after
// _equalize_root_cell, src/bdd_set.cpp:1472-1520). The root cell of the reference representation is canonical
// (minimal covering dyadic cell of height >= 0, because every mutating operation ends with minimize_height), so it
// is recomputed from the keys when cells have to be enumerated.
//
// Enumeration (reference: BDDTreeConstIterator / _compute_next_cell, src/bdd_set.cpp:800-891) yields the MAXIMAL
// dyadic cells contained in the set, left-first, with boxes obtained by bisecting the root cell.
#pragma once
#include <vector>
#include <stdint.h>
#include <functional>
#include "ab200_grid.cuh"

namespace ab200 {

struct GridDesc {
    int dim;
    double origin[AB_GRID_MAXN];
    double lengths[AB_GRID_MAXN];
};

struct CellBox { double b[2 * AB_GRID_MAXN]; };

class CellSet {
  public:
    CellSet() : loc(0) { grid.dim = 0; }
    CellSet(const GridDesc& g, const KeyLayout& k, int location) : grid(g), K(k), loc(location) { }

    GridDesc grid;
    KeyLayout K;
    int loc;
    std::vector<uint64_t> keys;      // sorted ascending, unique

    bool empty() const { return keys.empty(); }
    size_t leaf_count() const { return keys.size(); }

    void adjoin(const std::vector<uint64_t>& sorted_unique);
    void adjoin(const CellSet& o) { adjoin(o.keys); }
    void remove(const CellSet& o);
    void restrict(const CellSet& o);

    // adjoin_outer_approximation(Box,accuracy) (src/bdd_set.cpp:1548-1574), returns false if a cell fell outside
    // the key window.
    bool adjoin_outer_approximation(const double* box /* lo,hi per coordinate */);

    // A three-valued test on a cell box (lo,hi per coordinate): 0 false, 1 true, 2 indeterminate.
    typedef std::function<int(const double*)> BoxTest;
    // outer_restrict(SetCheckerInterface,accuracy) (src/bdd_set.cpp:1706-1721 -> _approximate_restrict :640-690, AS_OUTER):
    // descend from the minimal root; a subtree is dropped where the check is false, kept whole where it is true,
    // and kept at the leaves where it stays indeterminate.
    // to_bdd_depth: descend only as deep as the set's own BDD goes -- BDDTreeSet::outer_restrict(OpenSetInterface)
    // (src/bdd_set.cpp:1670-1686) stops at BDDTreeSet::depth() (:1171-1190), the last variable its BDD depends on, which is
    // above the accuracy when the finest levels of the set are merged (a box whose faces pair up sibling cells).
    void outer_restrict(const BoxTest& check, bool to_bdd_depth = false);
    // The same descent level by level: all boxes of one level are tested in one call (a device batch), then the undecided
    // ones are split. The test is pure, so the result is the depth-first one.
    typedef std::function<void(const std::vector<CellBox>&, std::vector<int>&)> BatchTest;
    void outer_restrict_batched(const BatchTest& check);
    // adjoin_outer_approximation(CompactSetInterface,accuracy) (src/bdd_set.cpp:1548-1574,455-493): grow the root over
    // `bbox`, then mark every leaf reached without meeting a cell the set is definitely disjoint from
    // (`disjoint` returns 1 only for "definitely disjoint"). Returns false if a cell fell outside the key window.
    bool adjoin_outer_approximation(const double* bbox, const BoxTest& disjoint);


    // Minimal root (height>=0) covering all cells; undefined for the empty set.
    Root minimal_root() const;
    // Maximal dyadic cells, left-first. If mince_levels >= 0, cells are split down to at least that many levels
    // below the root (BDDTreeSet::mince, src/bdd_set.cpp:1367-1369).
    void enumerate(std::vector<CellBox>& out, int mince_levels = -1) const;
    // BDDTreeSet::size(): number of maximal cells (src/bdd_set.cpp:117-148).
    size_t size() const { std::vector<CellBox> v; enumerate(v); return v.size(); }
    // bounding box of the set as the reference computes it (src/bdd_set.cpp:893-928): hull of the leaf boxes.
    bool bounding_box(double* box) const;
};

// Box queries against a fixed cell set (the tree of root-to-leaf paths is built once).
//   superset(box): BDDTreeSet::superset(Box) (src/bdd_set.cpp:1325-1343,340-383) -- the box lies in the root cell and every
//                  subtree it reaches (splitting the box with the cells, closed intervals) is full;
//   disjoint(box): BDDTreeSet::disjoint(Box) (src/bdd_set.cpp:1346-1361,386-415) -- no maximal full subtree whose closed cell
//                  meets the closed box.
class CellTree {
  public:
    explicit CellTree(const CellSet& cs);
    bool superset(const double* box) const;
    bool disjoint(const double* box) const;
    bool empty() const { return paths.empty(); }
  private:
    std::vector<unsigned __int128> paths;
    CellBox root_cell;
    int depth, dim, first_split;
};

} // namespace ab200
