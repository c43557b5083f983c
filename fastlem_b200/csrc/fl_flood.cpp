// fl_flood.cpp -- flood order of lake removal (static host prep, like the CSR itself).
//
// remove_lakes_from_stream_tree (reference src/lem/stream_tree.rs:175-243) floods the site graph from
// the outlets with a std::collections::BinaryHeap keyed on edge length only.  Which node is popped
// when depends on (graph, outlets) alone -- not on elevations, receivers or parameters -- so the pop
// sequence number T(i) of every node is a static property of the uploaded model.  The device uses T
// to connect every lake basin in parallel (fl_kernels.cuh, k_lake_min): the sequential flood connects
// lake L through the first (pop order of i, slot of j in adj(i)) pair with subroot[j] == L.
//
// T must reproduce the reference's order at exact key ties too (all outlets enter with key 0.0;
// equal edge lengths), so this is a faithful replay of Rust's BinaryHeap: push = append + sift_up,
// pop = swap last into the root + sift_down_to_bottom + sift_up, comparisons via `<=` on the
// reversed order of RidgeElement (stream_tree.rs:34-38).
#include "fl_flood.h"

#include <vector>

namespace {

struct Entry {
    double key;
    uint32_t node;
};

// max-heap on the reversed order: a "<=" b  <=>  a.key >= b.key
struct ReplayHeap {
    std::vector<Entry> h;

    static bool le(const Entry& a, const Entry& b) { return a.key >= b.key; }

    void rise(size_t floor, size_t at) {
        Entry moving = h[at];
        while (at > floor) {
            size_t up = (at - 1) >> 1;
            if (le(moving, h[up])) break;
            h[at] = h[up];
            at = up;
        }
        h[at] = moving;
    }

    void push(double key, uint32_t node) {
        h.push_back(Entry{key, node});
        rise(0, h.size() - 1);
    }

    // precondition: !h.empty()
    Entry pop() {
        Entry last = h.back();
        h.pop_back();
        if (h.empty()) return last;
        Entry top = h[0];
        // sink `last` from the root all the way to a leaf, always following the child that is not "<=" its sibling
        const size_t len = h.size();
        size_t at = 0, kid = 1;
        const size_t bound = len >= 2 ? len - 2 : 0;
        while (kid <= bound) {
            if (le(h[kid], h[kid + 1])) ++kid;
            h[at] = h[kid];
            at = kid;
            kid = 2 * at + 1;
        }
        if (kid == len - 1) {
            h[at] = h[kid];
            at = kid;
        }
        h[at] = last;
        rise(0, at);
        return top;
    }
};

}  // namespace

void fl_flood_rank(uint32_t n, const uint32_t* row_ptr, const uint32_t* col, const double* dist,
                   const uint32_t* outlets, uint32_t n_outlets, uint32_t* rank) {
    fl_flood_rank_prefix(n, row_ptr, col, dist, outlets, n_outlets, rank, 0xFFFFFFFFu);
}

uint32_t fl_flood_rank_prefix(uint32_t n, const uint32_t* row_ptr, const uint32_t* col, const double* dist,
                              const uint32_t* outlets, uint32_t n_outlets, uint32_t* rank, uint32_t stop_after) {
    for (uint32_t i = 0; i < n; ++i) rank[i] = FL_RANK_NONE;
    ReplayHeap heap;
    heap.h.reserve((size_t)n + 16);
    for (uint32_t k = 0; k < n_outlets; ++k) heap.push(0.0, outlets[k]);
    uint32_t seq = 0;
    while (!heap.h.empty() && seq < stop_after) {
        Entry e = heap.pop();
        uint32_t i = e.node;
        if (rank[i] != FL_RANK_NONE) continue;  // visited
        // every unvisited neighbour is pushed, in adjacency order (stream_tree.rs:204-236); the node itself
        // only counts as visited after its neighbours were scanned (:239)
        for (uint32_t s = row_ptr[i]; s < row_ptr[i + 1]; ++s) {
            uint32_t j = col[s];
            if (rank[j] == FL_RANK_NONE) heap.push(dist[s], j);
        }
        rank[i] = seq++;
    }
    return seq;
}
