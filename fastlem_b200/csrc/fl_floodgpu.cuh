// fl_floodgpu.cuh -- the flood order of lake removal, computed on the device.
//
// remove_lakes_from_stream_tree (reference src/lem/stream_tree.rs:175-243) pops sites from a binary heap keyed on
// edge length, starting from all outlets with key 0.0.  With lazy deletion that is Prim's algorithm on the site
// graph with the outlets contracted into one source S, so -- when all edge lengths are distinct and positive -- the
// pop order T of the other sites is a function of the MINIMUM SPANNING TREE only:
//
//   Root the MST at S; let w(v) be the length of v's parent edge.  Prim leaves the component "everything reachable
//   over edges lighter than w(v)" only through its lightest outgoing edge, hence (Kruskal reconstruction tree) for an
//   MST edge e = (a -> b): all of A_e (the sites reachable from a over lighter edges) is popped before any of B_e
//   (the same from b), and B_e lies inside b's subtree.  Let nga(v) be the nearest ancestor of v whose parent edge is
//   heavier than w(v) (S if none).  Then
//       B_e(v)   = the subtree of v in the forest of nga pointers,
//       |A_e(v)| = |{nga(v)}| + sum of |B| over the nga-siblings of v with a lighter parent edge
//                  (|{S}| = number of outlets),
//       T(v)     = |A_e(v)| + T(nga(v)),   T(S) = 0.
//   Everything is integer work: Boruvka rounds (64-bit atomicMin on the length bits), a frontier walk that roots the
//   tree, pointer walks for nga, a counting sweep for |B|, two radix sorts + scans for the sibling prefix, and
//   pointer jumping for the final sums.
//
// The outlets' own ranks depend on the heap's behaviour at equal keys (all 0.0) and have a closed form (k_flg_outlet_*,
// below): the first outlet, then the node-right-left pre-order of the array-embedded heap left by the first pop.  Its
// validity condition is checked on the device; when it fails the host replays just that prefix (fl_flood_rank_prefix).
// Graphs with equal or non-positive edge lengths fall back to the full host replay (fl_flood.cpp), which reproduces
// Rust's BinaryHeap at ties.
#pragma once
#include "fl_kernels.cuh"

#define FLG_KEY_NONE 0xFFFFFFFFFFFFFFFFull

struct FlFloodG {
    uint32_t n;    // sites; index n = the contracted source S
    uint32_t src;  // label of the source component (an outlet)
    const uint32_t* row_ptr;
    const uint32_t* col;
    const double* dist;
    const uint8_t* rev;
    const uint8_t* is_outlet;
    uint32_t* comp;               // n
    uint32_t* link;               // n
    unsigned long long* best;     // n
    uint32_t* pick;               // n: slot of the lightest outgoing edge of a component
    uint8_t* mst;                 // nnz: slot belongs to the spanning tree
    uint32_t* tptr;               // n+1: compact adjacency of the spanning tree (built after the Boruvka rounds)
    uint32_t* tcol;               // 2(n-1)
    unsigned long long* twb;      // 2(n-1): edge length bits
    uint32_t* par;                // n+1: parent in the rooted tree (n = S), FL_NONE = not reached
    unsigned long long* wbits;    // n+1: bits of the parent edge length
    uint32_t* nga;                // n+1
    uint32_t* cnt;                // n+1: nga children still to report
    uint32_t* size;               // n+1: |B|
    uint32_t* flags;              // [0] any component hooked, [1] next frontier size, [2] work left, [3] bad edge length,
                                  // [4..6] counters of the rooting walk, [7] closed form of the outlets' ranks not valid
};

__device__ __forceinline__ unsigned long long flg_bits(double d) { return (unsigned long long)__double_as_longlong(d); }

// undirected edge lengths (each edge once, from its lower endpoint) as sortable keys; flags[3] on a non-positive or
// non-finite length.  Slots that are not the lower endpoint get FLG_KEY_NONE (sorted last).
// Edges between two outlets take no part: they are self-loops of the contracted source.  In the heap flood they only
// ever produce entries for sites that are already in the heap with key 0.0 and are popped (visited) before any
// positive key, i.e. stale entries that are skipped when they surface -- so their lengths may tie freely.  This is
// the case of the reference's own `add_edge_sites` rim (equally spaced boundary sites, builder.rs:115-122) under an
// ocean mask (examples/terrain_generation_advanced.rs:178-182).
__global__ void __launch_bounds__(256) k_flg_edge_keys(FlFloodG g, unsigned long long* keys) {
    const uint32_t i = FL_TID;
    if (i >= g.n) return;
    const bool out_i = g.is_outlet[i] != 0;
    for (uint32_t s = g.row_ptr[i]; s < g.row_ptr[i + 1]; ++s) {
        const double d = g.dist[s];
        const uint32_t j = g.col[s];
        const bool inner = out_i && g.is_outlet[j] != 0;
        const bool bad = !(d > 0.0) || !(d < 1.7976931348623157e308);
        if (bad) g.flags[inner ? 7 : 3] = 1u;  // [7]: only the closed form of the outlets' ranks needs these positive
        keys[s] = (j > i && !inner) ? flg_bits(d) : FLG_KEY_NONE;
    }
}
__global__ void __launch_bounds__(256) k_flg_dup_check(uint32_t m, const unsigned long long* __restrict__ sorted,
                                                        uint32_t* flags) {
    const uint32_t i = FL_TID;
    if (i + 1u >= m) return;
    const unsigned long long a = sorted[i];
    if (a != FLG_KEY_NONE && a == sorted[i + 1u]) flags[3] = 1u;
}

// ------------------------------------------------------------------------------------------------
// The outlets' own ranks (stream_tree.rs:184-197 push every outlet with key 0.0; Rust's BinaryHeap decides their order).
// Equal keys never move in sift_up (`<=`), so after the pushes the heap array is the outlet list itself.  The first pop
// returns outlets[0]; sift_down_to_bottom follows the RIGHT child at equal keys, so the hole runs down the right spine
// 0, 2, 6, 14, ... and the last outlet lands at its end.  From then on every pop takes the root and promotes, level by
// level, the right child if both children have key 0.0, else the one that has (a positive key never beats 0.0) -- the
// 0.0 entries leave in node-right-left pre-order of that array, PROVIDED the element moved from the tail by each of
// these pops has a positive key (then it never rises above a 0.0 entry).  Sufficient: before the pop number k >= 1 at
// least k entries have been pushed (the array is then longer than the K - 1 slots the 0.0 entries live in).  Pushes of
// a pop = its neighbours not yet visited = non-outlets and outlets of a later (or its own) rank.
struct FlOutletPrefix {
    uint32_t count;       // K = number of outlets
    uint32_t right_steps; // moves of the first pop's hole along the right spine
    uint32_t final_left;  // 1: one more move into a last, single (left) child
    uint32_t spine_end;   // position reached by the right moves
    uint32_t last_pos;    // where the last outlet lands
};
inline FlOutletPrefix flg_outlet_prefix(uint32_t count) {
    FlOutletPrefix p;
    p.count = count; p.right_steps = 0u; p.final_left = 0u; p.spine_end = 0u; p.last_pos = 0u;
    if (count < 2u) return p;
    const unsigned long long len = (unsigned long long)count - 1ull, bound = len >= 2ull ? len - 2ull : 0ull;
    unsigned long long at = 0ull, kid = 1ull;
    while (kid <= bound) { at = kid + 1ull; kid = 2ull * at + 1ull; ++p.right_steps; }
    p.spine_end = (uint32_t)at;
    if (kid == len - 1ull) { p.final_left = 1u; at = kid; }
    p.last_pos = (uint32_t)at;
    return p;
}
// nodes of the subtree of r in the array-embedded complete binary tree of m nodes
__device__ __forceinline__ uint32_t flg_heap_subtree(uint32_t r, uint32_t m) {
    unsigned long long lo = r, hi = r, s = 0ull;
    while (lo < m) {
        s += (hi < m ? hi : (unsigned long long)m - 1ull) - lo + 1ull;
        lo = 2ull * lo + 1ull;
        hi = 2ull * hi + 2ull;
    }
    return (uint32_t)s;
}
// index of position p in the node-right-left pre-order of that tree
__device__ __forceinline__ uint32_t flg_nrl_index(uint32_t p, uint32_t m) {
    uint32_t before = 0u;
    while (p > 0u) {
        before += 1u;                                                           // the parent itself
        if ((p & 1u) && p + 1u < m) before += flg_heap_subtree(p + 1u, m);      // a left child: the right sibling's subtree
        p = (p - 1u) >> 1;
    }
    return before;
}
// rank of every outlet (outlet_rank: n words, preset to FL_NONE)
__global__ void __launch_bounds__(256) k_flg_outlet_rank(FlOutletPrefix f, const uint32_t* __restrict__ outlets,
                                                          uint32_t* outlet_rank) {
    const uint32_t a = FL_TID;
    if (a >= f.count) return;
    uint32_t r = 0u;
    if (a > 0u) {
        uint32_t pos = a;  // position in the array after the first pop
        if (a == f.count - 1u) pos = f.last_pos;
        else {
            const uint32_t d = 31u - (uint32_t)__clz((int)(a + 2u));  // a + 2 == 2^d: a lies on the right spine
            if (a + 2u == (1u << d) && d >= 2u && d <= f.right_steps + 1u) pos = (1u << (d - 1u)) - 2u;
            else if (f.final_left && a == 2u * f.spine_end + 1u) pos = f.spine_end;
        }
        r = 1u + flg_nrl_index(pos, f.count - 1u);
    }
    outlet_rank[outlets[a]] = r;
}
// pushes made while outlet number `rank` is popped
__global__ void __launch_bounds__(256) k_flg_outlet_pushes(FlFloodG g, uint32_t count, const uint32_t* __restrict__ outlets,
                                                            const uint32_t* __restrict__ outlet_rank, uint32_t* pushes) {
    const uint32_t a = FL_TID;
    if (a >= count) return;
    const uint32_t o = outlets[a], r = outlet_rank[o];
    uint32_t c = 0u;
    for (uint32_t s = g.row_ptr[o]; s < g.row_ptr[o + 1u]; ++s) {
        const uint32_t rj = outlet_rank[g.col[s]];
        if (rj == FL_NONE || rj >= r) ++c;
    }
    pushes[r] = c;
}
// before[k] = pushes of the pops 0 .. k-1
__global__ void __launch_bounds__(256) k_flg_outlet_check(uint32_t count, const uint32_t* __restrict__ before,
                                                           uint32_t* flags) {
    const uint32_t k = FL_TID;
    if (k == 0u || k >= count) return;
    if (before[k] < k) flags[7] = 1u;
}

__global__ void __launch_bounds__(256) k_flg_init(FlFloodG g) {
    const uint32_t v = FL_TID;
    if (v > g.n) return;
    g.par[v] = FL_NONE;
    g.nga[v] = FL_NONE;
    g.cnt[v] = 0u;
    g.size[v] = 1u;
    g.wbits[v] = 0ull;
    if (v == g.n) return;
    const uint32_t c = g.is_outlet[v] ? g.src : v;
    g.comp[v] = c;
    g.link[v] = c;
    g.best[v] = FLG_KEY_NONE;
}

// Boruvka round, step 1: lightest edge leaving each component
__global__ void __launch_bounds__(256) k_flg_min(FlFloodG g) {
    const uint32_t v = FL_TID;
    if (v >= g.n) return;
    const uint32_t cv = g.comp[v];
    unsigned long long m = FLG_KEY_NONE;
    for (uint32_t s = g.row_ptr[v]; s < g.row_ptr[v + 1]; ++s) {
        if (g.comp[g.col[s]] == cv) continue;
        const unsigned long long k = flg_bits(g.dist[s]);
        if (k < m) m = k;
    }
    if (m != FLG_KEY_NONE) atomicMin(&g.best[cv], m);
}
// step 2: the slot that carries it (lengths are distinct: exactly one slot inside the component)
__global__ void __launch_bounds__(256) k_flg_pick(FlFloodG g) {
    const uint32_t v = FL_TID;
    if (v >= g.n) return;
    const uint32_t cv = g.comp[v];
    const unsigned long long b = g.best[cv];
    if (b == FLG_KEY_NONE) return;
    for (uint32_t s = g.row_ptr[v]; s < g.row_ptr[v + 1]; ++s)
        if (flg_bits(g.dist[s]) == b && g.comp[g.col[s]] != cv) { g.pick[cv] = s; return; }
}
// step 3: hook each component onto the one its edge leads to; of two components that chose the same edge the
// smaller label stays a root.  The chosen edges join the spanning tree (both directions).
__global__ void __launch_bounds__(256) k_flg_hook(FlFloodG g) {
    const uint32_t r = FL_TID;
    if (r >= g.n) return;
    if (g.comp[r] != r) return;  // not a component label
    const unsigned long long b = g.best[r];
    if (b == FLG_KEY_NONE) return;
    const uint32_t s = g.pick[r];
    const uint32_t u = g.col[s];
    const uint32_t cu = g.comp[u];
    // owner of slot s: the site v with row_ptr[v] <= s < row_ptr[v+1]; reverse slot through rev
    g.mst[s] = 1u;
    g.mst[g.row_ptr[u] + g.rev[s]] = 1u;
    const bool mutual = g.best[cu] == b;
    if (!mutual || r > cu) g.link[r] = cu;
    g.flags[0] = 1u;
}
// step 4: new labels = roots of the hook forest; reset for the next round
__global__ void __launch_bounds__(256) k_flg_relabel(FlFloodG g) {
    const uint32_t v = FL_TID;
    if (v >= g.n) return;
    uint32_t r = g.comp[v];
    for (;;) {
        const uint32_t up = g.link[r];
        if (up == r) break;
        r = up;
    }
    g.comp[v] = r;
}
__global__ void __launch_bounds__(256) k_flg_reset(FlFloodG g) {
    const uint32_t v = FL_TID;
    if (v >= g.n) return;
    g.best[v] = FLG_KEY_NONE;
    g.link[v] = g.comp[v];  // every site now points at its (root) label; labels point at themselves
}

// rooting: outlets form level 0 (they ARE the source)
__global__ void __launch_bounds__(256) k_flg_root_init(FlFloodG g, uint32_t* frontier, uint32_t* cnt3) {
    const uint32_t v = FL_TID;
    if (v >= g.n) return;
    if (!g.is_outlet[v]) return;
    g.par[v] = g.n;
    frontier[atomicAdd(&cnt3[0], 1u)] = v;
}
// compact adjacency of the spanning tree, so that the level walk below touches tree edges only
__global__ void __launch_bounds__(256) k_flg_tree_degree(FlFloodG g, uint32_t* deg) {
    const uint32_t v = FL_TID;
    if (v > g.n) return;
    uint32_t d = 0;
    if (v < g.n)
        for (uint32_t s = g.row_ptr[v]; s < g.row_ptr[v + 1]; ++s) d += g.mst[s] ? 1u : 0u;
    deg[v] = d;
}
__global__ void __launch_bounds__(256) k_flg_tree_fill(FlFloodG g) {
    const uint32_t v = FL_TID;
    if (v >= g.n) return;
    uint32_t k = g.tptr[v];
    for (uint32_t s = g.row_ptr[v]; s < g.row_ptr[v + 1]; ++s)
        if (g.mst[s]) { g.tcol[k] = g.col[s]; g.twb[k] = flg_bits(g.dist[s]); ++k; }
}

// one frontier vertex: adopt the tree neighbours that have no parent yet
__device__ __forceinline__ void flg_expand(const FlFloodG& g, uint32_t v, uint32_t* next, uint32_t* next_count) {
    const uint32_t pv = g.is_outlet[v] ? g.n : v;  // children of an outlet hang under S
    const uint32_t k0 = g.tptr[v], k1 = g.tptr[v + 1];
    for (uint32_t kb = k0; kb < k1; kb += 4u) {  // four neighbours' loads in flight together
        uint32_t u[4], pu[4];
        unsigned long long w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            u[j] = FL_NONE; w[j] = 0ull;
            if (kb + (uint32_t)j < k1) { u[j] = g.tcol[kb + j]; w[j] = g.twb[kb + j]; }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) pu[j] = (u[j] != FL_NONE) ? g.par[u[j]] : 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (u[j] == FL_NONE || pu[j] != FL_NONE) continue;  // its own parent (or an outlet)
            g.par[u[j]] = pv;
            g.wbits[u[j]] = w[j];
            g.nga[u[j]] = pv;  // first candidate
            next[atomicAdd(next_count, 1u)] = u[j];
        }
    }
}

// The spanning tree is thousands of levels deep with ~100 sites per level: one CTA walks ALL levels, a block barrier
// between them, instead of one launch per level.
__global__ void __launch_bounds__(1024) k_flg_root_walk(FlFloodG g, uint32_t* fa, uint32_t* fb, uint32_t* cnt3) {
#ifdef FL_EMU
    if (FL_TID != 0u) return;
    uint32_t count = cnt3[0];
    uint32_t* cur = fa;
    uint32_t* nxt = fb;
    while (count) {
        uint32_t produced = 0;
        for (uint32_t t = 0; t < count; ++t) flg_expand(g, cur[t], nxt, &produced);
        uint32_t* tmp = cur; cur = nxt; nxt = tmp;
        count = produced;
    }
#else
    __shared__ uint32_t s_count, s_next;
    uint32_t* cur = fa;
    uint32_t* nxt = fb;
    if (threadIdx.x == 0) { s_count = cnt3[0]; s_next = 0u; }
    __syncthreads();
    for (;;) {
        const uint32_t count = s_count;
        if (count == 0u) break;
        for (uint32_t t = threadIdx.x; t < count; t += blockDim.x) flg_expand(g, cur[t], nxt, &s_next);
        __syncthreads();
        if (threadIdx.x == 0) { s_count = s_next; s_next = 0u; }
        uint32_t* tmp = cur; cur = nxt; nxt = tmp;
        __syncthreads();
    }
#endif
}

// nearest ancestor with a heavier parent edge.  Invariant: every ancestor strictly between v and nga[v] has a parent
// edge lighter than w(v); any value another thread has stored for an ancestor satisfies the same for that
// ancestor, so reading it mid-way is safe.
__global__ void __launch_bounds__(256) k_flg_nga(FlFloodG g, uint32_t max_hops) {
    const uint32_t v = FL_TID;
    if (v >= g.n) return;
    if (g.is_outlet[v] || g.par[v] == FL_NONE) return;
    const unsigned long long wv = g.wbits[v];
    volatile uint32_t* nga = g.nga;
    uint32_t c = nga[v];
    uint32_t hops = 0;
    while (c != g.n && g.wbits[c] < wv) {
        if (hops++ == max_hops) { g.flags[2] = 1u; break; }
        c = nga[c];
    }
    nga[v] = c;
}

// |B|: children counts, then a counting sweep from the leaves of the nga forest (integer sums: order-free)
__global__ void __launch_bounds__(256) k_flg_count_children(FlFloodG g) {
    const uint32_t v = FL_TID;
    if (v >= g.n) return;
    if (g.is_outlet[v] || g.par[v] == FL_NONE) return;
    atomicAdd(&g.cnt[g.nga[v]], 1u);
}
// sites with children get a count biased by one, so that "no children" (0) stays distinguishable from "all children
// have reported" (1) while the sweep runs
__global__ void __launch_bounds__(256) k_flg_bias_counts(FlFloodG g) {
    const uint32_t v = FL_TID;
    if (v > g.n) return;
    if (g.cnt[v] != 0u) g.cnt[v] += 1u;
}
__global__ void __launch_bounds__(256) k_flg_sizes(FlFloodG g) {
    uint32_t v = FL_TID;
    if (v >= g.n) return;
    if (g.is_outlet[v] || g.par[v] == FL_NONE) return;
    if (g.cnt[v] != 0u) return;  // not a leaf of the nga forest
    uint32_t sz = 1u;
    for (;;) {
        const uint32_t p = g.nga[v];
        if (p == g.n) return;
        atomicAdd(&g.size[p], sz);
        __threadfence();
        if (atomicSub(&g.cnt[p], 1u) != 2u) return;  // somebody else arrives last (counts are biased by one)
        __threadfence();  // acquire side: the other children's additions to size[p]
        sz = atomicAdd(&g.size[p], 0u);
        v = p;
    }
}

// sibling prefix: sort by (nga parent, parent edge length); sizes in that order; group starts
__global__ void __launch_bounds__(256) k_flg_sort_keys(FlFloodG g, unsigned long long* wkey, uint32_t* ids) {
    const uint32_t v = FL_TID;
    if (v >= g.n) return;
    ids[v] = v;
    wkey[v] = (g.is_outlet[v] || g.par[v] == FL_NONE) ? FLG_KEY_NONE : g.wbits[v];
}
__global__ void __launch_bounds__(256) k_flg_parent_keys(FlFloodG g, const uint32_t* __restrict__ ids,
                                                          unsigned long long* pkey) {
    const uint32_t i = FL_TID;
    if (i >= g.n) return;
    const uint32_t v = ids[i];
    pkey[i] = (g.is_outlet[v] || g.par[v] == FL_NONE) ? 0xFFFFFFFFull : (unsigned long long)g.nga[v];
}
__global__ void __launch_bounds__(256) k_flg_group(FlFloodG g, const unsigned long long* __restrict__ pkey_sorted,
                                                    const uint32_t* __restrict__ ids_sorted,
                                                    unsigned long long* sz_sorted, uint32_t* gstart) {
    const uint32_t i = FL_TID;
    if (i >= g.n) return;
    sz_sorted[i] = (pkey_sorted[i] == 0xFFFFFFFFull) ? 0ull : (unsigned long long)g.size[ids_sorted[i]];
    gstart[i] = (i == 0u || pkey_sorted[i] != pkey_sorted[i - 1u]) ? i : 0u;
}
// pd[v] = (nga[v], |A_e(v)|) for the final pointer jumping; sites outside the flood point at themselves
__global__ void __launch_bounds__(256) k_flg_terms(FlFloodG g, uint32_t n_outlets,
                                                    const unsigned long long* __restrict__ pkey_sorted,
                                                    const uint32_t* __restrict__ ids_sorted,
                                                    const unsigned long long* __restrict__ scan,
                                                    const uint32_t* __restrict__ gstart_scan,
                                                    unsigned long long* pd) {
    const uint32_t i = FL_TID;
    if (i > g.n) return;
    if (i == g.n) { pd[g.n] = (unsigned long long)g.n; return; }
    const uint32_t v = ids_sorted[i];
    if (pkey_sorted[i] == 0xFFFFFFFFull) { pd[v] = (unsigned long long)v; return; }
    const uint32_t p = (uint32_t)pkey_sorted[i];
    const unsigned long long lighter = scan[i] - scan[gstart_scan[i]];
    const unsigned long long a = (p == g.n ? (unsigned long long)n_outlets : 1ull) + lighter;
    pd[v] = (unsigned long long)p | (a << 32);
}
__global__ void __launch_bounds__(256) k_flg_finish(FlFloodG g, const unsigned long long* __restrict__ pd,
                                                     const uint32_t* __restrict__ outlet_rank, uint32_t* rank) {
    const uint32_t v = FL_TID;
    if (v >= g.n) return;
    if (g.is_outlet[v]) { rank[v] = outlet_rank[v]; return; }
    rank[v] = (g.par[v] == FL_NONE) ? FL_NONE : (uint32_t)(pd[v] >> 32);
}
