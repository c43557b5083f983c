// fl_paths.cuh -- path-decomposed tree sweeps ("sweep" >= 1).
//
// Why: both tree sweeps of an iteration are chains of NON-associative double additions whose order the
// reference fixes (generator.rs:154-174): A[j] = ((a_j + A[c_k]) + ...) + A[c_1] bottom-up and
// rt_i = rt_recv + t_i top-down.  Bit-exact results therefore need depth-many dependent additions
// (~800 .. 5000 levels at 1M sites).  Level-synchronous launches pay a kernel launch per level; here the
// forest is cut into PATHS (a node continues into one chosen "heavy" child), every path is laid out
// contiguously in memory, and one thread walks a whole path with the running value in a register.  Paths
// nest only O(log N) deep when the heavy child is the one with the largest drainage area, so a sweep is
// ~20 rounds instead of ~1000 levels.  The additions themselves are performed in exactly the reference order.
//
// The layout is a renumbering of the sites: internal id = position.  Position q+1 is the heavy child of q
// inside a path.  Adjacency ORDER per row is preserved by the renumbering, so every tie-break of the
// reference (first slot wins, children in reverse slot order) is unaffected.
#pragma once
#include "fl_kernels.cuh"

#define FL_LONG_PATH 32u  // paths at least this long go to the warp kernels ("sweep" = 2)
#ifndef FL_EMU
#define FL_FULL 0xFFFFFFFFu
__device__ __forceinline__ double fl_shfl(double v, int src) { return __shfl_sync(FL_FULL, v, src); }
#endif

// ------------------------------------------------------------------------------------------------
// static per-graph table: rev[s] = slot of i inside adj(col[s]) (first match), 255 if >= 255 / absent
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rev_slots(uint32_t n, const uint32_t* __restrict__ row_ptr,
                                                    const uint32_t* __restrict__ col, uint8_t* __restrict__ rev) {
    uint32_t i = FL_TID;
    if (i >= n) return;
    const uint32_t s1 = row_ptr[i + 1];
    for (uint32_t s = row_ptr[i]; s < s1; ++s) {
        const uint32_t j = col[s];
        const uint32_t t0 = row_ptr[j], t1 = row_ptr[j + 1];
        uint32_t r = 255;
        for (uint32_t t = t0; t < t1; ++t)
            if (col[t] == i) { r = (t - t0) < 255u ? (t - t0) : 255u; break; }
        rev[s] = (uint8_t)r;
    }
}

// ------------------------------------------------------------------------------------------------
// K1 (stream_tree.rs:109-137) + child registration: the chosen receiver gets bit `slot of i in its row`
// set in its child mask, so parents can later enumerate children in adjacency order without rescanning
// their neighbours.  Rows longer than 32 ignore the mask and rescan (fl_children_rev).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_receivers_mask(uint32_t n, const uint32_t* __restrict__ row_ptr,
                                                         const uint32_t* __restrict__ col,
                                                         const double* __restrict__ dist,
                                                         const uint8_t* __restrict__ rev,
                                                         const double* __restrict__ elev,
                                                         const uint8_t* __restrict__ is_outlet,
                                                         uint32_t* __restrict__ recv, double* __restrict__ drecv,
                                                         uint32_t* cmask, uint32_t* __restrict__ flags) {
    uint32_t i = FL_TID;
    if (i >= n) return;
    uint32_t best = i, best_s = FL_NONE;
    double best_d = 1.0;
    if (!is_outlet[i]) {
        const double ei = elev[i];
        double steepest = 0.0;
        const uint32_t s1 = row_ptr[i + 1];
        for (uint32_t s = row_ptr[i]; s < s1; ++s) {
            const uint32_t j = col[s];
            const double ej = elev[j];
            if (ei > ej) {
                const double d = dist[s];
                const double slope = (ei - ej) / d;
                if (slope > steepest) {
                    steepest = slope;
                    best = j;
                    best_d = d;
                    best_s = s;
                }
            }
        }
        if (best == i) atomicOr(&flags[FL_FLAG_LAKE], 1u);
    }
    recv[i] = best;
    drecv[i] = best_d;
    if (best_s != FL_NONE) {
        const uint32_t r = rev[best_s];
        if (r < 32u) atomicOr(&cmask[best], 1u << r);
    }
}

// child masks from scratch (after lake removal rewrote receivers)
__global__ void __launch_bounds__(256) k_childmask(uint32_t n, const uint32_t* __restrict__ row_ptr,
                                                    const uint32_t* __restrict__ col, const uint8_t* __restrict__ rev,
                                                    const uint32_t* __restrict__ recv, uint32_t* cmask) {
    uint32_t i = FL_TID;
    if (i >= n) return;
    const uint32_t p = recv[i];
    if (p == i) return;
    const uint32_t s1 = row_ptr[i + 1];
    for (uint32_t s = row_ptr[i]; s < s1; ++s)
        if (col[s] == p) {
            const uint32_t r = rev[s];
            if (r < 32u) atomicOr(&cmask[p], 1u << r);
            return;
        }
}

// children of q in REVERSE adjacency order (the order generator.rs:154-159 adds them in)
template <class F>
__device__ __forceinline__ void fl_children_rev(uint32_t q, const uint32_t* __restrict__ row_ptr,
                                                const uint32_t* __restrict__ col, const uint32_t* __restrict__ recv,
                                                const uint32_t* __restrict__ cmask, F&& f) {
    const uint32_t s0 = row_ptr[q];
    const uint32_t deg = row_ptr[q + 1] - s0;
    if (deg <= 32u) {
        uint32_t m = cmask[q];
        while (m) {
            const uint32_t b = 31u - (uint32_t)__clz((int)m);
            m ^= 1u << b;
            f(col[s0 + b]);
        }
    } else {
        for (uint32_t s = deg; s-- > 0;) {
            const uint32_t c = col[s0 + s];
            if (c != q && recv[c] == q) f(c);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// layout rebuild, step 1: heavy child = child with the largest weight (previous drainage area); the
// first such child in reverse adjacency order wins.  Any choice is CORRECT; it only shapes the paths.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_heavy(uint32_t n, const uint32_t* __restrict__ row_ptr,
                                                const uint32_t* __restrict__ col, const uint32_t* __restrict__ recv,
                                                const uint32_t* __restrict__ cmask, const double* __restrict__ weight,
                                                uint32_t* __restrict__ heavy) {
    uint32_t q = FL_TID;
    if (q >= n) return;
    uint32_t best = FL_NONE;
    double bw = 0.0;
    fl_children_rev(q, row_ptr, col, recv, cmask, [&](uint32_t c) {
        const double w = weight[c];
        if (best == FL_NONE || w > bw) { best = c; bw = w; }
    });
    heavy[q] = best;
}

// step 2: chain pointers for list ranking along paths: up = parent if I am its heavy child, else myself
__global__ void __launch_bounds__(256) k_chain_init(uint32_t n, const uint32_t* __restrict__ recv,
                                                     const uint32_t* __restrict__ heavy,
                                                     unsigned long long* __restrict__ pd) {
    uint32_t c = FL_TID;
    if (c >= n) return;
    const uint32_t p = recv[c];
    const bool chained = (p != c) && (heavy[p] == c);
    pd[c] = chained ? ((unsigned long long)p | (1ull << 32)) : (unsigned long long)c;
}

// step 3: path length, written by the path's last node (no heavy child): plen[head] = pos + 1
__global__ void __launch_bounds__(256) k_path_len(uint32_t n, const uint32_t* __restrict__ heavy,
                                                   const unsigned long long* __restrict__ pd,
                                                   uint32_t* __restrict__ plen) {
    uint32_t c = FL_TID;
    if (c >= n) return;
    if (heavy[c] != FL_NONE) return;
    const unsigned long long a = pd[c];
    plen[(uint32_t)a] = (uint32_t)(a >> 32) + 1u;
}

// step 4: nesting level of each path = number of path switches between its head and the tree root.
// Jump table over path heads: head -> head of the parent's path.
__global__ void __launch_bounds__(256) k_nest_init(uint32_t n, const uint32_t* __restrict__ recv,
                                                    const unsigned long long* __restrict__ pd,
                                                    unsigned long long* __restrict__ pd2) {
    uint32_t q = FL_TID;
    if (q >= n) return;
    const uint32_t head = (uint32_t)pd[q];
    if (head != q) { pd2[q] = (unsigned long long)q; return; }  // not a head: idle entry
    const uint32_t p = recv[q];
    if (p == q) { pd2[q] = (unsigned long long)q; return; }  // tree root: level 0
    pd2[q] = (unsigned long long)(uint32_t)pd[p] | (1ull << 32);
}

// step 5: sort keys for path heads (FL_NONE for everything else): the nesting level, or with `split`
// (level * 2 + (path shorter than FL_LONG_PATH)) so that the long paths of a level come first.
__global__ void __launch_bounds__(256) k_path_keys(uint32_t n, const unsigned long long* __restrict__ pd,
                                                    const unsigned long long* __restrict__ pd2,
                                                    const uint32_t* __restrict__ plen, int split,
                                                    uint32_t* __restrict__ keys, uint32_t* __restrict__ ids,
                                                    uint32_t* flags) {
    uint32_t q = FL_TID;
    if (q >= n) return;
    ids[q] = q;
    if ((uint32_t)pd[q] == q) {
        const uint32_t level = (uint32_t)(pd2[q] >> 32);
        const uint32_t key = split ? (level * 2u + (plen[q] < FL_LONG_PATH ? 1u : 0u)) : level;
        keys[q] = key;
        if (key > 0) atomicMax(&flags[FL_FLAG_MAXDEPTH], key);
    } else {
        keys[q] = FL_NONE;
    }
}

// step 6: per sorted path k: its length and the inverse map head -> k
__global__ void __launch_bounds__(256) k_path_gather(uint32_t npaths, const uint32_t* __restrict__ hlist,
                                                      const uint32_t* __restrict__ plen,
                                                      uint32_t* __restrict__ len_sorted,
                                                      uint32_t* __restrict__ hrank) {
    uint32_t k = FL_TID;
    if (k >= npaths) return;
    const uint32_t h = hlist[k];
    len_sorted[k] = plen[h];
    hrank[h] = k;
}

// step 7: new position of every node = start of its path + position in the path
__global__ void __launch_bounds__(256) k_newpos(uint32_t n, const unsigned long long* __restrict__ pd,
                                                 const uint32_t* __restrict__ hrank,
                                                 const uint32_t* __restrict__ starts, uint32_t* __restrict__ newpos) {
    uint32_t q = FL_TID;
    if (q >= n) return;
    const unsigned long long a = pd[q];
    newpos[q] = starts[hrank[(uint32_t)a]] + (uint32_t)(a >> 32);
}

// ------------------------------------------------------------------------------------------------
// renumbering
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_deg_scatter(uint32_t n, const uint32_t* __restrict__ row_ptr,
                                                      const uint32_t* __restrict__ newpos,
                                                      uint32_t* __restrict__ deg_new) {
    uint32_t q = FL_TID;
    if (q >= n) return;
    deg_new[newpos[q]] = row_ptr[q + 1] - row_ptr[q];
}

__global__ void __launch_bounds__(256) k_permute_rows(uint32_t n, const uint32_t* __restrict__ row_ptr,
                                                       const uint32_t* __restrict__ col,
                                                       const double* __restrict__ dist,
                                                       const uint8_t* __restrict__ rev,
                                                       const uint32_t* __restrict__ newpos,
                                                       const uint32_t* __restrict__ row_ptr_new,
                                                       uint32_t* __restrict__ col_new, double* __restrict__ dist_new,
                                                       uint8_t* __restrict__ rev_new) {
    uint32_t q = FL_TID;
    if (q >= n) return;
    const uint32_t s0 = row_ptr[q], deg = row_ptr[q + 1] - s0;
    const uint32_t t0 = row_ptr_new[newpos[q]];
    for (uint32_t k = 0; k < deg; ++k) {
        col_new[t0 + k] = newpos[col[s0 + k]];
        dist_new[t0 + k] = dist[s0 + k];
        rev_new[t0 + k] = rev[s0 + k];
    }
}

struct FlNodeArrays {
    // f64 per-node arrays (tan may be null)
    const double *areas, *erod, *uplift, *tan, *elev, *drecv;
    double *areas_n, *erod_n, *uplift_n, *tan_n, *elev_n, *drecv_n;
    // u32
    const uint32_t *recv, *cmask, *rank, *orig_of;
    uint32_t *recv_n, *cmask_n, *rank_n, *orig_of_n;
    const uint8_t* is_outlet;
    uint8_t* is_outlet_n;
};

__global__ void __launch_bounds__(256) k_permute_nodes(uint32_t n, const uint32_t* __restrict__ newpos,
                                                        FlNodeArrays a) {
    uint32_t q = FL_TID;
    if (q >= n) return;
    const uint32_t p = newpos[q];
    a.areas_n[p] = a.areas[q];
    a.erod_n[p] = a.erod[q];
    a.uplift_n[p] = a.uplift[q];
    if (a.tan) a.tan_n[p] = a.tan[q];
    a.elev_n[p] = a.elev[q];
    a.drecv_n[p] = a.drecv[q];
    a.recv_n[p] = newpos[a.recv[q]];
    a.cmask_n[p] = a.cmask[q];
    if (a.rank) a.rank_n[p] = a.rank[q];
    a.orig_of_n[p] = a.orig_of[q];
    a.is_outlet_n[p] = a.is_outlet[q];
}

__global__ void __launch_bounds__(256) k_rank_inverse(uint32_t n, const uint32_t* __restrict__ rank,
                                                       uint32_t* __restrict__ rank_to_node) {
    uint32_t q = FL_TID;
    if (q >= n) return;
    const uint32_t t = rank[q];
    if (t != FL_NONE) rank_to_node[t] = q;
}

__global__ void __launch_bounds__(256) k_iota(uint32_t n, uint32_t* __restrict__ a) {
    uint32_t q = FL_TID;
    if (q < n) a[q] = q;
}

// out[orig_of[q]] = in[q]  (results back in the caller's numbering)
__global__ void __launch_bounds__(256) k_unpermute_f64(uint32_t n, const uint32_t* __restrict__ orig_of,
                                                        const double* __restrict__ in, double* __restrict__ out) {
    uint32_t q = FL_TID;
    if (q < n) out[orig_of[q]] = in[q];
}

// same for arrays whose VALUES are node ids (receivers, labels); FL_NONE stays
__global__ void __launch_bounds__(256) k_unpermute_ids(uint32_t n, const uint32_t* __restrict__ orig_of,
                                                        const uint32_t* __restrict__ in, uint32_t* __restrict__ out) {
    uint32_t q = FL_TID;
    if (q >= n) return;
    const uint32_t v = in[q];
    out[orig_of[q]] = (v == FL_NONE) ? FL_NONE : orig_of[v];
}

__global__ void __launch_bounds__(256) k_unpermute_u32(uint32_t n, const uint32_t* __restrict__ orig_of,
                                                        const uint32_t* __restrict__ in, uint32_t* __restrict__ out) {
    uint32_t q = FL_TID;
    if (q < n) out[orig_of[q]] = in[q];
}

__global__ void __launch_bounds__(256) k_gather_u32(uint32_t n, const uint32_t* __restrict__ orig_of,
                                                     const uint32_t* __restrict__ src, uint32_t* __restrict__ dst) {
    uint32_t q = FL_TID;
    if (q < n) dst[q] = src[orig_of[q]];
}

// debug stages: label / depth of the final forest; depth FL_NONE where the root is not an outlet
__global__ void __launch_bounds__(256) k_labels_depth(uint32_t n, const unsigned long long* __restrict__ pd,
                                                       const uint8_t* __restrict__ is_outlet,
                                                       uint32_t* __restrict__ label, uint32_t* __restrict__ depth) {
    uint32_t q = FL_TID;
    if (q >= n) return;
    const unsigned long long a = pd[q];
    label[q] = (uint32_t)a;
    depth[q] = is_outlet[(uint32_t)a] ? (uint32_t)(a >> 32) : FL_NONE;
}

// debug stages: value where the site is visited, the loop's initial value (areas / 0.0) elsewhere
__global__ void __launch_bounds__(256) k_stage_value(uint32_t n, const uint32_t* __restrict__ depth,
                                                      const double* __restrict__ val,
                                                      const double* __restrict__ fallback, double* __restrict__ out) {
    uint32_t q = FL_TID;
    if (q >= n) return;
    out[q] = depth[q] != FL_NONE ? val[q] : (fallback ? fallback[q] : 0.0);
}

// in[int] <- src[orig]  (parameters / initial elevations into the current numbering)
__global__ void __launch_bounds__(256) k_gather_f64(uint32_t n, const uint32_t* __restrict__ orig_of,
                                                     const double* __restrict__ src, double* __restrict__ dst) {
    uint32_t q = FL_TID;
    if (q < n) dst[q] = src[orig_of[q]];
}

// ------------------------------------------------------------------------------------------------
// K4 on paths (generator.rs:154-159): one thread per path, bottom (tail) to top (head); `x` carries the
// heavy child's finished area in a register.  Children other than q+1 are heads of deeper-nested paths,
// finished in an earlier round.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_area_paths(uint32_t begin, uint32_t count,
                                                     const uint32_t* __restrict__ seg_head,
                                                     const uint32_t* __restrict__ seg_len,
                                                     const uint32_t* __restrict__ row_ptr,
                                                     const uint32_t* __restrict__ col,
                                                     const uint32_t* __restrict__ recv,
                                                     const uint32_t* __restrict__ cmask,
                                                     const double* __restrict__ areas, double* A) {
    uint32_t t = FL_TID;
    if (t >= count) return;
    const uint32_t h = seg_head[begin + t];
    const uint32_t last = h + seg_len[begin + t] - 1u;
    double x = 0.0;
    for (uint32_t q = last;; --q) {
        double a = areas[q];
        const bool has_chain = q < last;
        fl_children_rev(q, row_ptr, col, recv, cmask, [&](uint32_t c) {
            a += (has_chain && c == q + 1u) ? x : A[c];
        });
        A[q] = a;
        x = a;
        if (q == h) break;
    }
}

// ------------------------------------------------------------------------------------------------
// K5 on paths (generator.rs:162-203): one thread per path, head to tail; response time and the new
// elevation of the receiver travel down the path in registers.
//   rt_i = 0.0 + (rt_recv + 1.0 / (k_i * sqrt(A_i)) * d_i);  z = e_out + u_i * max(rt_i - rt_out, 0.0)
// root_of[q] = root of q's tree (FL_NONE if that root is not an outlet: such trees are never visited).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_elev_paths(uint32_t begin, uint32_t count,
                                                     const uint32_t* __restrict__ seg_head,
                                                     const uint32_t* __restrict__ seg_len,
                                                     const uint32_t* __restrict__ recv,
                                                     const double* __restrict__ drecv, const double* __restrict__ A,
                                                     const double* __restrict__ erod,
                                                     const double* __restrict__ uplift,
                                                     const double* __restrict__ tan_slope,
                                                     const uint8_t* __restrict__ is_outlet, double* elev, double* rt,
                                                     uint32_t* root_of, uint32_t* __restrict__ flags) {
    uint32_t t = FL_TID;
    if (t >= count) return;
    const uint32_t h = seg_head[begin + t];
    const uint32_t last = h + seg_len[begin + t] - 1u;
    const uint32_t p = recv[h];
    const bool is_root = (p == h);
    uint32_t root;
    double rt_prev, z_prev, e_out, rt_out;
    if (is_root) {
        root = is_outlet[h] ? h : FL_NONE;
        rt_prev = 0.0;
        z_prev = elev[h];  // has_edge(i,i) is false -> the clamp compares with the site's own old elevation
        e_out = elev[h];
        rt_out = 0.0;  // set below from the root's own response time
    } else {
        root = root_of[p];
        rt_prev = rt[p];
        z_prev = elev[p];  // receiver already holds its NEW elevation
        e_out = root != FL_NONE ? elev[root] : 0.0;
        rt_out = root != FL_NONE ? rt[root] : 0.0;
    }
    if (root == FL_NONE) {
        for (uint32_t q = h; q <= last; ++q) root_of[q] = FL_NONE;
        return;
    }
    bool changed = false;
    for (uint32_t q = h; q <= last; ++q) {
        const double d = drecv[q];
        const double celerity = erod[q] * sqrt(A[q]);
        const double rti = 0.0 + (rt_prev + 1.0 / celerity * d);
        if (is_root && q == h) rt_out = rti;
        double z = e_out + uplift[q] * fmax(rti - rt_out, 0.0);
        if (tan_slope) {
            const double ms = tan_slope[q];
            if (ms == ms) {
                const double slope = (z - z_prev) / d;
                if (slope > ms) z = z_prev + ms * d;
            }
        }
        changed |= (z != elev[q]);
        if (is_root && q == h) e_out = z;  // later sites read elevations[outlet] after the outlet's own update
        elev[q] = z;
        rt[q] = rti;
        root_of[q] = root;
        rt_prev = rti;
        z_prev = z;
    }
    if (changed) flags[FL_FLAG_CHANGED] = 1u;
}

// ================================================================================================
// Long paths: one WARP per path ("sweep" = 2).  A single thread walking a 1000-site path pays a DRAM/L2
// round trip per site (~3 us); here the 32 lanes fetch a 32-site chunk of the path together (coalesced,
// the path is contiguous), every lane prepares its own site's order-independent part, and only the truly
// sequential double additions run as a 32-step chain over warp shuffles -- in exactly the reference order.
// ================================================================================================

#ifndef FL_EMU
// t-th child (0-based) of q AFTER the chain child q+1 in reverse adjacency order (rare overflow path)
__device__ double fl_post_child_value(uint32_t q, uint32_t t, const uint32_t* __restrict__ row_ptr,
                                      const uint32_t* __restrict__ col, const uint32_t* __restrict__ recv,
                                      const uint32_t* __restrict__ cmask, const double* A) {
    bool seen = false;
    uint32_t k = 0;
    double out = 0.0;
    fl_children_rev(q, row_ptr, col, recv, cmask, [&](uint32_t c) {
        if (c == q + 1u) { seen = true; return; }
        if (seen) { if (k == t) out = A[c]; ++k; }
    });
    return out;
}

// K4, warp per long path.  For site q with chain child q+1:
//   A[q] = ((pre_q + A[q+1]) + post_1) + post_2 ...   pre_q = a_q + children before the chain child (reverse
//   adjacency order), post_* = children after it.  pre/post are gathered by lane (parallel), the chain
//   x -> A[q] runs over the lanes in order.
__global__ void __launch_bounds__(128) k_area_paths_warp(uint32_t begin, uint32_t count,
                                                          const uint32_t* __restrict__ seg_head,
                                                          const uint32_t* __restrict__ seg_len,
                                                          const uint32_t* __restrict__ row_ptr,
                                                          const uint32_t* __restrict__ col,
                                                          const uint32_t* __restrict__ recv,
                                                          const uint32_t* __restrict__ cmask,
                                                          const double* __restrict__ areas, double* A) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= count) return;
    const uint32_t h = seg_head[begin + w];
    const uint32_t last = h + seg_len[begin + w] - 1u;
    double x = 0.0;
    for (long long top = last; top >= (long long)h; top -= 32) {
        const long long qq = top - lane;
        const bool active = qq >= (long long)h;
        const uint32_t q = (uint32_t)qq;
        const bool has_chain = active && q < last;
        double pre = 0.0, p1 = 0.0, p2 = 0.0;
        uint32_t np = 0;
        if (active) {
            pre = areas[q];
            bool seen = false;
            fl_children_rev(q, row_ptr, col, recv, cmask, [&](uint32_t c) {
                if (has_chain && c == q + 1u) { seen = true; return; }
                const double v = A[c];
                if (!seen) pre += v;
                else { if (np == 0) p1 = v; else if (np == 1) p2 = v; ++np; }
            });
        }
        const int nact = (int)((top - (long long)h + 1) < 32 ? (top - (long long)h + 1) : 32);
        double mine = 0.0;
        for (int k = 0; k < nact; ++k) {
            const double pre_k = fl_shfl(pre, k);
            const double p1_k = fl_shfl(p1, k);
            const double p2_k = fl_shfl(p2, k);
            const uint32_t np_k = __shfl_sync(FL_FULL, np, k);
            const int hc_k = __shfl_sync(FL_FULL, (int)has_chain, k);
            double y = hc_k ? (pre_k + x) : pre_k;
            if (np_k >= 1) y += p1_k;
            if (np_k >= 2) y += p2_k;
            for (uint32_t t = 2; t < np_k; ++t) {  // more than two children after the chain child: rare
                double v = 0.0;
                if (lane == k) v = fl_post_child_value(q, t, row_ptr, col, recv, cmask, A);
                y += fl_shfl(v, k);
            }
            x = y;
            if (lane == k) mine = y;
        }
        if (active) A[q] = mine;
    }
}

// K5, warp per long path (generator.rs:162-203).  t_q = 1/(k_q*sqrt(A_q))*d_q is per-lane work; the response
// time chain rt_q = 0.0 + (rt_{q-1} + t_q) and (only with max_slope) the clamp chain run over the lanes.
__global__ void __launch_bounds__(128) k_elev_paths_warp(uint32_t begin, uint32_t count,
                                                          const uint32_t* __restrict__ seg_head,
                                                          const uint32_t* __restrict__ seg_len,
                                                          const uint32_t* __restrict__ recv,
                                                          const double* __restrict__ drecv,
                                                          const double* __restrict__ A,
                                                          const double* __restrict__ erod,
                                                          const double* __restrict__ uplift,
                                                          const double* __restrict__ tan_slope,
                                                          const uint8_t* __restrict__ is_outlet, double* elev,
                                                          double* rt, uint32_t* root_of,
                                                          uint32_t* __restrict__ flags) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= count) return;
    const uint32_t h = seg_head[begin + w];
    const uint32_t last = h + seg_len[begin + w] - 1u;
    const uint32_t p = recv[h];
    uint32_t root, first = h;
    double rt_prev, z_prev, e_out, rt_out;
    bool changed = false;
    if (p == h) {  // tree root: handled by every lane redundantly (uniform), written by lane 0
        if (!is_outlet[h]) {
            for (uint32_t q = h + lane; q <= last; q += 32) root_of[q] = FL_NONE;
            return;
        }
        root = h;
        const double e_old = elev[h];
        const double rti = 0.0 + (0.0 + 1.0 / (erod[h] * sqrt(A[h])) * drecv[h]);
        double z = e_old + uplift[h] * fmax(rti - rti, 0.0);
        if (tan_slope) {
            const double ms = tan_slope[h];
            if (ms == ms) {
                const double d = drecv[h];
                const double slope = (z - e_old) / d;
                if (slope > ms) z = e_old + ms * d;
            }
        }
        changed = (z != e_old);
        __syncwarp();
        if (lane == 0) { elev[h] = z; rt[h] = rti; root_of[h] = h; }
        rt_prev = rti; z_prev = z; e_out = z; rt_out = rti;
        first = h + 1u;
    } else {
        root = root_of[p];
        if (root == FL_NONE) {
            for (uint32_t q = h + lane; q <= last; q += 32) root_of[q] = FL_NONE;
            return;
        }
        rt_prev = rt[p];
        z_prev = elev[p];
        e_out = elev[root];
        rt_out = rt[root];
    }
    for (uint32_t base = first; base <= last; base += 32) {
        const uint32_t q = base + lane;
        const bool active = q <= last;
        double t = 0.0, u = 0.0, zold = 0.0, d = 1.0, ms = 0.0;
        if (active) {
            d = drecv[q];
            t = 1.0 / (erod[q] * sqrt(A[q])) * d;
            u = uplift[q];
            zold = elev[q];
            if (tan_slope) ms = tan_slope[q];
        }
        const int nact = (int)((last - base + 1u) < 32u ? (last - base + 1u) : 32u);
        double my_rt = 0.0;
        for (int k = 0; k < nact; ++k) {
            const double t_k = fl_shfl(t, k);
            rt_prev = 0.0 + (rt_prev + t_k);
            if (lane == k) my_rt = rt_prev;
        }
        double z = e_out + u * fmax(my_rt - rt_out, 0.0);
        if (tan_slope) {
            double my_z = z;
            for (int k = 0; k < nact; ++k) {
                double z_k = fl_shfl(z, k);
                const double ms_k = fl_shfl(ms, k);
                const double d_k = fl_shfl(d, k);
                if (ms_k == ms_k) {
                    const double slope = (z_k - z_prev) / d_k;
                    if (slope > ms_k) z_k = z_prev + ms_k * d_k;
                }
                z_prev = z_k;
                if (lane == k) my_z = z_k;
            }
            z = my_z;
        }
        if (active) {
            changed |= (z != zold);
            elev[q] = z;
            rt[q] = my_rt;
            root_of[q] = root;
        }
    }
    if (changed) flags[FL_FLAG_CHANGED] = 1u;
}
#endif  // !FL_EMU

// ================================================================================================
// "sweep" = 3: dataflow sweeps on DYNAMIC segments.
//
// A segment is a maximal run of positions q, q+1, ... with recv[q+1] == q (a chain that is contiguous in the
// current numbering).  Right after a layout rebuild the segments are exactly the heavy paths; when receivers
// change in later iterations a chain simply breaks into shorter segments -- nothing else has to be updated,
// so the numbering can be kept for many iterations and is rebuilt only when it has degraded.
//
// K4 (drainage area) runs as ONE launch without any level structure:
//   * every leaf starts a scan that climbs its segment with the running area in a register;
//   * a scan that finishes a segment head h (A[h] final) reports to the parent site p = recv[h];
//     the LAST child of p to report ("last arriver") gathers p's non-chain children in reverse adjacency
//     order into   pre  = a_p + (children before the chain child)   and   post1, post2 (children after it),
//     then either resumes the scan that is waiting at p or leaves the values for the scan still to come;
//   * nobody ever spins: a thread either continues with work that is ready or exits.
// The additions are the reference's, in the reference's order (generator.rs:154-159).
// It also yields, per segment head, the nesting height (longest chain of segment hand-offs below it), which
// orders the top-down sweep exactly for the CURRENT forest.
// ================================================================================================
#define FL_ST_COUNT_MASK 0x00FFFFFFu
#define FL_ST_NP_SHIFT 24
#define FL_ST_NP_MASK 0x0F000000u
#define FL_ST_PRE_READY 0x40000000u
#define FL_ST_SCAN_ARRIVED 0x80000000u

#ifdef FL_EMU
template <class T> __device__ __forceinline__ T fl_ld_cg(const T* p) { return *p; }
__device__ __forceinline__ uint32_t fl_ld_acquire(const uint32_t* p) { return *p; }
#else
template <class T> __device__ __forceinline__ T fl_ld_cg(const T* p) { return __ldcg(p); }
// acquire load: later loads of this thread (pre / posts) are ordered after it
__device__ __forceinline__ uint32_t fl_ld_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#endif
// relaxed (pipelinable) load of a flag word; order later loads with one __threadfence() per batch
#ifdef FL_EMU
__device__ __forceinline__ uint32_t fl_ld_relaxed(const uint32_t* p) { return *p; }
#else
__device__ __forceinline__ uint32_t fl_ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#endif
#define FL_BATCH 8

struct FlFlow {
    uint32_t n;
    const uint32_t* row_ptr;
    const uint32_t* col;
    const uint32_t* recv;
    const uint32_t* cmask;
    const double* areas;
    double* A;
    uint32_t* state;  // zeroed before the launch
    double* pre;
    double* post1;
    double* post2;
    double* xbuf;     // running area handed over at a waiting site
    uint32_t* hbuf;   // running nesting height handed over with it
    uint32_t* hgt;    // out: nesting height for segment heads, FL_NONE elsewhere
    uint32_t* hpre;   // max height over the non-chain children of a site (+1), written with pre
    uint32_t* flags;
    uint32_t* parked;    // sites where a long scan was parked for the warp-level pass
    uint32_t* counters;  // [0] = number of parked scans, [1] = next one to take
    uint32_t park_after; // a thread parks its scan after climbing this many sites in a row (0 = never)
};

// non-chain children of p, reverse adjacency order -> pre / posts; returns np (15 = more than two posts)
__device__ __forceinline__ uint32_t fl_gather_lights(const FlFlow& f, uint32_t p, bool has_chain, double& pre,
                                                     double& p1, double& p2, uint32_t& hmax) {
    pre = f.areas[p];
    p1 = 0.0; p2 = 0.0;
    uint32_t np = 0;
    bool seen = false;
    hmax = 0;
    const uint32_t s0 = f.row_ptr[p];
    uint32_t m = f.cmask[p];
    while (m) {
        const uint32_t b = 31u - (uint32_t)__clz((int)m);
        m ^= 1u << b;
        const uint32_t c = f.col[s0 + b];
        if (has_chain && c == p + 1u) { seen = true; continue; }
        const double v = fl_ld_cg(&f.A[c]);
        const uint32_t hc = fl_ld_cg(&f.hgt[c]) + 1u;
        if (hc > hmax) hmax = hc;
        if (!seen) pre += v;
        else { if (np == 0) p1 = v; else if (np == 1) p2 = v; ++np; }
    }
    return np > 2 ? 15u : np;
}

// slow path for np == 15: add every post child in order
__device__ double fl_add_posts(const FlFlow& f, uint32_t p, double y) {
    bool seen = false;
    const uint32_t s0 = f.row_ptr[p];
    uint32_t m = f.cmask[p];
    while (m) {
        const uint32_t b = 31u - (uint32_t)__clz((int)m);
        m ^= 1u << b;
        const uint32_t c = f.col[s0 + b];
        if (c == p + 1u) { seen = true; continue; }
        if (seen) y += fl_ld_cg(&f.A[c]);
    }
    return y;
}

// One thread follows one flow: climb, report to the parent, possibly take over as last arriver, ...
__device__ void fl_flow_thread(const FlFlow& f, uint32_t cur, double x, uint32_t hrun, bool has_chain, bool may_park) {
    // x: finished area of the chain child (valid when has_chain); hrun: nesting height along this segment
    uint32_t climbed = 0;  // sites climbed through the fast path without a break
    // `resume`: pre/posts of `cur` already in registers (we are the last arriver continuing the scan)
    bool resume = false;
    double pre = 0.0, p1 = 0.0, p2 = 0.0;
    uint32_t np = 0, hp = 0;
    for (;;) {
        double y = 0.0;
        uint32_t p = FL_NONE;
        bool at_head = false;
        // ---- fast path: up to FL_BATCH consecutive sites of the chain, every load issued up front ----
        if (!resume) {
            const uint32_t nb = cur + 1u < (uint32_t)FL_BATCH ? cur + 1u : (uint32_t)FL_BATCH;
            uint32_t cm[FL_BATCH], rc[FL_BATCH], st[FL_BATCH], hq[FL_BATCH];
            double ar[FL_BATCH], pr[FL_BATCH], q1[FL_BATCH], q2[FL_BATCH];
#pragma unroll
            for (int k = 0; k < FL_BATCH; ++k) {
                cm[k] = 0u; rc[k] = FL_NONE; ar[k] = 0.0;
                if ((uint32_t)k < nb) { cm[k] = f.cmask[cur - k]; rc[k] = f.recv[cur - k]; ar[k] = f.areas[cur - k]; }
            }
#pragma unroll
            for (int k = 0; k < FL_BATCH; ++k) {
                st[k] = 0u;
                const uint32_t nl = (uint32_t)__popc(cm[k]) - ((k > 0 || has_chain) ? 1u : 0u);
                if ((uint32_t)k < nb && cm[k] != 0u && nl > 0u) st[k] = fl_ld_relaxed(&f.state[cur - k]);
            }
            {
                uint32_t any = 0u;
#pragma unroll
                for (int k = 0; k < FL_BATCH; ++k) any |= st[k];
                if (any & FL_ST_PRE_READY) __threadfence();  // acquire: pre/posts are read after the flags
            }
#pragma unroll
            for (int k = 0; k < FL_BATCH; ++k) {
                pr[k] = 0.0; q1[k] = 0.0; q2[k] = 0.0; hq[k] = 0u;
                if (st[k] & FL_ST_PRE_READY) {
                    const uint32_t npk = (st[k] & FL_ST_NP_MASK) >> FL_ST_NP_SHIFT;
                    pr[k] = fl_ld_cg(&f.pre[cur - k]);
                    hq[k] = fl_ld_cg(&f.hpre[cur - k]);
                    if (npk >= 1u && npk != 15u) q1[k] = fl_ld_cg(&f.post1[cur - k]);
                    if (npk >= 2u && npk != 15u) q2[k] = fl_ld_cg(&f.post2[cur - k]);
                }
            }
            uint32_t done = 0;  // sites of the batch finished and climbed past
#pragma unroll
            for (int k = 0; k < FL_BATCH; ++k) {
                if (!at_head && done == (uint32_t)k && (uint32_t)k < nb) {
                    const uint32_t idx = cur - k;
                    const bool hc = (k > 0) || has_chain;
                    const uint32_t nl = (uint32_t)__popc(cm[k]) - (hc ? 1u : 0u);
                    bool ok = true;
                    if (nl == 0u) {
                        y = hc ? (ar[k] + x) : ar[k];
                    } else if (st[k] & FL_ST_PRE_READY) {
                        const uint32_t npk = (st[k] & FL_ST_NP_MASK) >> FL_ST_NP_SHIFT;
                        y = hc ? (pr[k] + x) : pr[k];
                        if (npk == 15u) y = fl_add_posts(f, idx, y);
                        else {
                            if (npk >= 1u) y += q1[k];
                            if (npk >= 2u) y += q2[k];
                        }
                        if (hq[k] > hrun) hrun = hq[k];
                    } else {
                        ok = false;  // children of idx still running: take the hand-off path below
                    }
                    if (ok) {
                        f.A[idx] = y;
                        if (idx > 0u && rc[k] == idx - 1u) {
                            f.hgt[idx] = FL_NONE;
                            x = y;
                            done = (uint32_t)k + 1u;
                        } else {
                            at_head = true;
                            p = rc[k];
                        }
                    }
                }
            }
            if (done > 0u) has_chain = true;
            cur -= done;
            if (!at_head && done == nb) {  // whole batch climbed
                climbed += done;
                if (may_park && f.park_after != 0u && climbed >= f.park_after) {
                    // a long chain: leave it to the warp-level pass (k_area_flow_long)
                    f.xbuf[cur] = x;
                    f.hbuf[cur] = hrun;
                    f.parked[atomicAdd(&f.counters[0], 1u)] = cur;
                    return;
                }
                continue;  // next batch
            }
            climbed = 0;
        }
        if (!at_head) {
            // ---- general path for one site: wait for / take over from its children ----
            const uint32_t m = f.cmask[cur];
            const uint32_t nlight = (uint32_t)__popc(m) - (has_chain ? 1u : 0u);
            if (nlight == 0u) {
                y = has_chain ? (f.areas[cur] + x) : f.areas[cur];
            } else {
                if (!resume) {
                    uint32_t s = fl_ld_acquire(&f.state[cur]);
                    if (!(s & FL_ST_PRE_READY)) {
                        f.xbuf[cur] = x;
                        f.hbuf[cur] = hrun;
                        __threadfence();
                        s = atomicOr(&f.state[cur], FL_ST_SCAN_ARRIVED);
                        if (!(s & FL_ST_PRE_READY)) return;  // the last arriver of `cur` takes over
                    }
                    __threadfence();
                    np = (s & FL_ST_NP_MASK) >> FL_ST_NP_SHIFT;
                    pre = fl_ld_cg(&f.pre[cur]);
                    hp = fl_ld_cg(&f.hpre[cur]);
                    if (np >= 1u && np != 15u) p1 = fl_ld_cg(&f.post1[cur]);
                    if (np >= 2u && np != 15u) p2 = fl_ld_cg(&f.post2[cur]);
                }
                resume = false;
                y = has_chain ? (pre + x) : pre;
                if (np == 15u) y = fl_add_posts(f, cur, y);
                else {
                    if (np >= 1u) y += p1;
                    if (np >= 2u) y += p2;
                }
                if (hp > hrun) hrun = hp;
            }
            f.A[cur] = y;
            p = f.recv[cur];
            if (cur > 0u && p == cur - 1u) {  // chained: climb
                f.hgt[cur] = FL_NONE;
                x = y;
                has_chain = true;
                cur = cur - 1u;
                continue;
            }
        }
        // ---- `cur` is a segment head with final area y; p = recv[cur] ----
        f.hgt[cur] = hrun;
        if (p == cur) {  // tree root: its segment has the largest nesting height of the tree
            if (hrun > 0u) atomicMax(&f.flags[FL_FLAG_MAXDEPTH], hrun);
            return;
        }
        __threadfence();  // publish A[cur], hgt[cur]
        const uint32_t arrived = (atomicAdd(&f.state[p], 1u) & FL_ST_COUNT_MASK) + 1u;
        const bool p_has_chain = (p + 1u < f.n) && (f.recv[p + 1u] == p);
        const uint32_t p_lights = (uint32_t)__popc(f.cmask[p]) - (p_has_chain ? 1u : 0u);
        if (arrived < p_lights) return;
        // last arriver at p: gather p's non-chain children
        __threadfence();
        np = fl_gather_lights(f, p, p_has_chain, pre, p1, p2, hp);
        if (!p_has_chain) {  // p ends its segment: nobody scans into it, start the scan here
            cur = p; has_chain = false; x = 0.0; hrun = 0; resume = true;
            continue;
        }
        f.pre[p] = pre;
        f.hpre[p] = hp;
        if (np >= 1u && np != 15u) f.post1[p] = p1;
        if (np >= 2u && np != 15u) f.post2[p] = p2;
        __threadfence();
        const uint32_t old = atomicOr(&f.state[p], FL_ST_PRE_READY | (np << FL_ST_NP_SHIFT));
        if (!(old & FL_ST_SCAN_ARRIVED)) return;  // the scan below p has not arrived yet; it will pick these up
        __threadfence();
        x = fl_ld_cg(&f.xbuf[p]);
        hrun = fl_ld_cg(&f.hbuf[p]);
        cur = p; has_chain = true; resume = true;
    }
}

// pass 1: every leaf starts a thread-level flow
__global__ void __launch_bounds__(256) k_area_flow(FlFlow f) {
    const uint32_t q0 = FL_TID;
    if (q0 >= f.n) return;
    if (f.cmask[q0] != 0u) return;  // only leaves (no children at all) start a scan
    fl_flow_thread(f, q0, 0.0, 0u, false, true);
}

#ifndef FL_EMU
__device__ __forceinline__ uint32_t fl_warp_max(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t w = __shfl_xor_sync(FL_FULL, v, o);
        v = w > v ? w : v;
    }
    return v;
}

// One WARP follows one flow.  All lanes hold identical copies of the flow state; a 32-site window of the
// chain is fetched by the lanes together, the additions run as a chain over shuffles.
__device__ void fl_flow_warp(const FlFlow& f, uint32_t cur, double x, uint32_t hrun, bool has_chain) {
    const int lane = threadIdx.x & 31;
    bool resume = false;
    double pre = 0.0, p1 = 0.0, p2 = 0.0;
    uint32_t np = 0, hp = 0;
    for (;;) {
        double y = 0.0;
        uint32_t p = FL_NONE;
        bool at_head = false;
        if (!resume) {
            const long long li = (long long)cur - lane;
            const bool valid = li >= 0;
            const uint32_t idx = (uint32_t)li;
            uint32_t cm = 0u, rc = FL_NONE;
            double ar = 0.0;
            if (valid) { cm = f.cmask[idx]; rc = f.recv[idx]; ar = f.areas[idx]; }
            const uint32_t rc_up = __shfl_up_sync(FL_FULL, rc, 1);
            const bool link = valid && (lane == 0 || rc_up == idx);
            const uint32_t linkmask = __ballot_sync(FL_FULL, link);
            const uint32_t nchain = linkmask == FL_FULL ? 32u : (uint32_t)__ffs((int)~linkmask) - 1u;
            const bool hc = (lane > 0) || has_chain;
            const bool inwin = (uint32_t)lane < nchain;
            const uint32_t nl = inwin ? (uint32_t)__popc(cm) - (hc ? 1u : 0u) : 0u;
            uint32_t st = 0u;
            if (inwin && nl > 0u) st = fl_ld_relaxed(&f.state[idx]);
            const bool ready = inwin && (nl == 0u || (st & FL_ST_PRE_READY));
            const uint32_t readymask = __ballot_sync(FL_FULL, ready);
            const uint32_t nproc = readymask == FL_FULL ? 32u : (uint32_t)__ffs((int)~readymask) - 1u;
            const bool mine_lit = (uint32_t)lane < nproc && nl > 0u;
            if (__ballot_sync(FL_FULL, mine_lit)) __threadfence();  // acquire before reading pre/posts
            double b = ar, q1 = 0.0, q2 = 0.0;
            uint32_t npk = 0u, hq = 0u;
            if (mine_lit) {
                npk = (st & FL_ST_NP_MASK) >> FL_ST_NP_SHIFT;
                b = fl_ld_cg(&f.pre[idx]);
                hq = fl_ld_cg(&f.hpre[idx]);
                if (npk >= 1u && npk != 15u) q1 = fl_ld_cg(&f.post1[idx]);
                if (npk >= 2u && npk != 15u) q2 = fl_ld_cg(&f.post2[idx]);
            }
            double mine = 0.0;
            for (uint32_t k = 0; k < nproc; ++k) {
                const double bk = fl_shfl(b, (int)k);
                const double q1k = fl_shfl(q1, (int)k);
                const double q2k = fl_shfl(q2, (int)k);
                const uint32_t npk_k = __shfl_sync(FL_FULL, npk, (int)k);
                double yy = ((k > 0u) || has_chain) ? (bk + x) : bk;
                if (npk_k == 15u) {
                    double v = 0.0;
                    if ((uint32_t)lane == k) v = fl_add_posts(f, idx, yy);
                    __syncwarp();
                    yy = fl_shfl(v, (int)k);
                } else {
                    if (npk_k >= 1u) yy += q1k;
                    if (npk_k >= 2u) yy += q2k;
                }
                x = yy;
                if ((uint32_t)lane == k) mine = yy;
            }
            const uint32_t hw = fl_warp_max((uint32_t)lane < nproc ? hq : 0u);
            if (hw > hrun) hrun = hw;
            const bool climbs = valid && idx > 0u && rc == idx - 1u;
            if ((uint32_t)lane < nproc) {
                f.A[idx] = mine;
                if (climbs) f.hgt[idx] = FL_NONE;
            }
            if (nproc > 0u) {
                const int lastl = (int)nproc - 1;
                const int last_climbs = __shfl_sync(FL_FULL, (int)climbs, lastl);
                if (!last_climbs) {
                    at_head = true;
                    y = fl_shfl(mine, lastl);
                    p = __shfl_sync(FL_FULL, rc, lastl);
                    cur -= (uint32_t)lastl;
                } else {
                    has_chain = true;
                    cur -= nproc;
                    if (nproc == 32u) continue;  // whole window climbed
                }
            }
        }
        if (!at_head) {
            // general path for the single site `cur` (uniform across the warp; lane 0 does the side effects)
            const uint32_t nlight = (uint32_t)__popc(f.cmask[cur]) - (has_chain ? 1u : 0u);
            if (nlight == 0u) {
                y = has_chain ? (f.areas[cur] + x) : f.areas[cur];
            } else {
                if (!resume) {
                    uint32_t sv = 0u;
                    if (lane == 0) {
                        sv = fl_ld_acquire(&f.state[cur]);
                        if (!(sv & FL_ST_PRE_READY)) {
                            f.xbuf[cur] = x;
                            f.hbuf[cur] = hrun;
                            __threadfence();
                            sv = atomicOr(&f.state[cur], FL_ST_SCAN_ARRIVED);
                        }
                    }
                    sv = __shfl_sync(FL_FULL, sv, 0);
                    if (!(sv & FL_ST_PRE_READY)) return;  // the last arriver of `cur` takes over
                    __threadfence();
                    np = (sv & FL_ST_NP_MASK) >> FL_ST_NP_SHIFT;
                    pre = fl_ld_cg(&f.pre[cur]);
                    hp = fl_ld_cg(&f.hpre[cur]);
                    if (np >= 1u && np != 15u) p1 = fl_ld_cg(&f.post1[cur]);
                    if (np >= 2u && np != 15u) p2 = fl_ld_cg(&f.post2[cur]);
                }
                resume = false;
                y = has_chain ? (pre + x) : pre;
                if (np == 15u) y = fl_add_posts(f, cur, y);
                else {
                    if (np >= 1u) y += p1;
                    if (np >= 2u) y += p2;
                }
                if (hp > hrun) hrun = hp;
            }
            if (lane == 0) f.A[cur] = y;
            p = f.recv[cur];
            if (cur > 0u && p == cur - 1u) {
                if (lane == 0) f.hgt[cur] = FL_NONE;
                x = y;
                has_chain = true;
                cur = cur - 1u;
                continue;
            }
        }
        // `cur` is a segment head with final area y
        if (lane == 0) f.hgt[cur] = hrun;
        if (p == cur) {
            if (lane == 0 && hrun > 0u) atomicMax(&f.flags[FL_FLAG_MAXDEPTH], hrun);
            return;
        }
        uint32_t arrived = 0u;
        if (lane == 0) {
            __threadfence();
            arrived = (atomicAdd(&f.state[p], 1u) & FL_ST_COUNT_MASK) + 1u;
        }
        arrived = __shfl_sync(FL_FULL, arrived, 0);
        const bool p_has_chain = (p + 1u < f.n) && (f.recv[p + 1u] == p);
        const uint32_t p_lights = (uint32_t)__popc(f.cmask[p]) - (p_has_chain ? 1u : 0u);
        if (arrived < p_lights) return;
        __threadfence();
        np = fl_gather_lights(f, p, p_has_chain, pre, p1, p2, hp);  // uniform addresses: every lane, same values
        if (!p_has_chain) {
            cur = p; has_chain = false; x = 0.0; hrun = 0; resume = true;
            continue;
        }
        uint32_t old = 0u;
        if (lane == 0) {
            f.pre[p] = pre;
            f.hpre[p] = hp;
            if (np >= 1u && np != 15u) f.post1[p] = p1;
            if (np >= 2u && np != 15u) f.post2[p] = p2;
            __threadfence();
            old = atomicOr(&f.state[p], FL_ST_PRE_READY | (np << FL_ST_NP_SHIFT));
        }
        old = __shfl_sync(FL_FULL, old, 0);
        if (!(old & FL_ST_SCAN_ARRIVED)) return;
        __threadfence();
        x = fl_ld_cg(&f.xbuf[p]);
        hrun = fl_ld_cg(&f.hbuf[p]);
        cur = p; has_chain = true; resume = true;
    }
}
#endif

// pass 2: the parked (long) scans.  Persistent: every warp (emulation: thread) takes parked scans until none is left.
__global__ void __launch_bounds__(256) k_area_flow_long(FlFlow f) {
#ifdef FL_EMU
    for (;;) {
        const uint32_t i = atomicAdd(&f.counters[1], 1u);
        if (i >= f.counters[0]) return;
        const uint32_t cur = f.parked[i];
        fl_flow_thread(f, cur, f.xbuf[cur], f.hbuf[cur], true, false);
    }
#else
    const int lane = threadIdx.x & 31;
    for (;;) {
        uint32_t i = 0u;
        if (lane == 0) i = atomicAdd(&f.counters[1], 1u);
        i = __shfl_sync(FL_FULL, i, 0);
        if (i >= fl_ld_cg(&f.counters[0])) return;
        const uint32_t cur = f.parked[i];
        fl_flow_warp(f, cur, fl_ld_cg(&f.xbuf[cur]), fl_ld_cg(&f.hbuf[cur]), true);
    }
#endif
}

// K5 on dynamic segments: one thread per segment head, walks while recv[q+1] == q.  Sites are taken in
// batches (4, then 8): all loads of a batch are issued together and the per-site celerity terms
// t = 1/(k*sqrt(A))*d are computed side by side, so that only the running additions are serial.
struct FlElev {
    uint32_t n;
    const uint32_t* recv;
    const double* drecv;
    const double* tcel;  // 1/(k*sqrt(A))*d per site (k_celerity_term)
    const double* A;
    const double* erod;
    const double* uplift;
    const double* tan_slope;  // may be null
    const uint8_t* is_outlet;
    double* elev;
    double* rt;
    uint32_t* root_of;
    uint32_t* flags;
};

// generator.rs:172-173: celerity = k_i * A_i^0.5;  term = 1.0 / celerity * d_i  (fully parallel; the
// division and square root stay out of the serial scans)
__global__ void __launch_bounds__(256) k_celerity_term(uint32_t n, const double* __restrict__ erod,
                                                        const double* __restrict__ A,
                                                        const double* __restrict__ drecv, double* __restrict__ tcel) {
    uint32_t q = FL_TID;
    if (q >= n) return;
    const double celerity = erod[q] * sqrt(A[q]);
    tcel[q] = 1.0 / celerity * drecv[q];
}

template <int B>
__device__ __forceinline__ bool fl_elev_batch(const FlElev& e, uint32_t& q, uint32_t h, bool is_root, uint32_t root,
                                              double& rt_prev, double& z_prev, double& e_out, double& rt_out,
                                              bool& changed) {
    // returns true when the segment ended inside this batch
    const uint32_t nb = e.n - q < (uint32_t)B ? e.n - q : (uint32_t)B;
    double d[B], t[B], up[B], eo[B], ms[B];
    uint32_t nx[B];
#pragma unroll
    for (int k = 0; k < B; ++k) {
        d[k] = 1.0; t[k] = 1.0; up[k] = 0.0; eo[k] = 0.0; ms[k] = 0.0; nx[k] = FL_NONE;
        if ((uint32_t)k < nb) {
            const uint32_t i = q + k;
            if (e.tan_slope) d[k] = e.drecv[i];
            t[k] = e.tcel[i];
            up[k] = e.uplift[i];
            eo[k] = e.elev[i];
            if (e.tan_slope) ms[k] = e.tan_slope[i];
            nx[k] = (i + 1u < e.n) ? e.recv[i + 1u] : FL_NONE;
        }
    }
    bool ended = false;
#pragma unroll
    for (int k = 0; k < B; ++k) {
        if (!ended && (uint32_t)k < nb) {
            const uint32_t i = q + k;
            const double rti = 0.0 + (rt_prev + t[k]);
            if (is_root && i == h) rt_out = rti;
            double z = e_out + up[k] * fmax(rti - rt_out, 0.0);
            if (e.tan_slope) {
                if (ms[k] == ms[k]) {
                    const double slope = (z - z_prev) / d[k];
                    if (slope > ms[k]) z = z_prev + ms[k] * d[k];
                }
            }
            changed |= (z != eo[k]);
            if (is_root && i == h) e_out = z;
            e.elev[i] = z;
            e.rt[i] = rti;
            e.root_of[i] = root;
            rt_prev = rti;
            z_prev = z;
            if (nx[k] != i) ended = true;
        }
    }
    q += nb;
    return ended || q >= e.n;
}

#ifndef FL_EMU
// rest of a long segment, walked by the whole warp: 32-site windows, shuffle chains.  Returns "changed".
__device__ bool fl_elev_warp(const FlElev& e, uint32_t q, uint32_t root, double rt_prev, double z_prev, double e_out,
                             double rt_out) {
    const int lane = threadIdx.x & 31;
    bool changed = false;
    for (;;) {
        const uint32_t i = q + (uint32_t)lane;
        const bool valid = i < e.n;
        double t = 0.0, up = 0.0, eold = 0.0, ms = 0.0, d = 1.0;
        uint32_t nx = FL_NONE;
        if (valid) {
            t = e.tcel[i];
            up = e.uplift[i];
            eold = e.elev[i];
            nx = (i + 1u < e.n) ? e.recv[i + 1u] : FL_NONE;
            if (e.tan_slope) { ms = e.tan_slope[i]; d = e.drecv[i]; }
        }
        const uint32_t endmask = __ballot_sync(FL_FULL, !valid || nx != i);
        uint32_t nproc = 32u;
        if (endmask) {
            const int el = __ffs((int)endmask) - 1;
            const int el_valid = __shfl_sync(FL_FULL, (int)valid, el);
            nproc = (uint32_t)el + (el_valid ? 1u : 0u);
        }
        double my_rt = 0.0;
        for (uint32_t k = 0; k < nproc; ++k) {
            const double tk = fl_shfl(t, (int)k);
            rt_prev = 0.0 + (rt_prev + tk);
            if ((uint32_t)lane == k) my_rt = rt_prev;
        }
        double z = e_out + up * fmax(my_rt - rt_out, 0.0);
        if (e.tan_slope) {
            double my_z = z;
            for (uint32_t k = 0; k < nproc; ++k) {
                double zk = fl_shfl(z, (int)k);
                const double msk = fl_shfl(ms, (int)k);
                const double dk = fl_shfl(d, (int)k);
                if (msk == msk) {
                    const double slope = (zk - z_prev) / dk;
                    if (slope > msk) zk = z_prev + msk * dk;
                }
                z_prev = zk;
                if ((uint32_t)lane == k) my_z = zk;
            }
            z = my_z;
        }
        if ((uint32_t)lane < nproc) {
            changed |= (z != eold);
            e.elev[i] = z;
            e.rt[i] = my_rt;
            e.root_of[i] = root;
        }
        q += nproc;
        if (endmask) break;
    }
    return __ballot_sync(FL_FULL, changed) != 0u;
}
#endif

// One thread per segment head for the first 12 sites; segments that go on are finished by the whole warp,
// one after the other (GPU) -- or by the same thread (emulation).
__global__ void __launch_bounds__(128) k_elev_flow(uint32_t begin, uint32_t count,
                                                    const uint32_t* __restrict__ heads, FlElev e) {
    const uint32_t t = FL_TID;
    const bool active = t < count;
    bool changed = false, longseg = false;
    uint32_t q = 0, root = FL_NONE;
    double rt_prev = 0.0, z_prev = 0.0, e_out = 0.0, rt_out = 0.0;
    if (active) {
        const uint32_t h = heads[begin + t];
        const uint32_t p = e.recv[h];
        const bool is_root = (p == h);
        if (is_root) {
            root = e.is_outlet[h] ? h : FL_NONE;
            rt_prev = 0.0;
            z_prev = e.elev[h];  // has_edge(i,i) is false: the clamp compares with the site's own old elevation
            e_out = e.elev[h];
            rt_out = 0.0;
        } else {
            root = e.root_of[p];
            rt_prev = e.rt[p];
            z_prev = e.elev[p];  // the receiver already holds its NEW elevation
            e_out = root != FL_NONE ? e.elev[root] : 0.0;
            rt_out = root != FL_NONE ? e.rt[root] : 0.0;
        }
        if (root == FL_NONE) {  // tree without outlet: never visited (generator.rs:149)
            for (uint32_t r = h;; ++r) {
                e.root_of[r] = FL_NONE;
                if (r + 1u >= e.n || e.recv[r + 1u] != r) break;
            }
        } else {
            q = h;
            bool ended = fl_elev_batch<4>(e, q, h, is_root, root, rt_prev, z_prev, e_out, rt_out, changed);
            if (!ended) ended = fl_elev_batch<8>(e, q, h, is_root, root, rt_prev, z_prev, e_out, rt_out, changed);
            longseg = !ended;
#ifdef FL_EMU
            if (longseg) while (!fl_elev_batch<8>(e, q, h, false, root, rt_prev, z_prev, e_out, rt_out, changed)) {}
#endif
        }
    }
#ifndef FL_EMU
    uint32_t todo = __ballot_sync(FL_FULL, longseg);
    const int lane = threadIdx.x & 31;
    while (todo) {
        const int src = __ffs((int)todo) - 1;
        todo &= todo - 1u;
        const uint32_t q_s = __shfl_sync(FL_FULL, q, src);
        const uint32_t root_s = __shfl_sync(FL_FULL, root, src);
        const double rtp_s = fl_shfl(rt_prev, src);
        const double zp_s = fl_shfl(z_prev, src);
        const double eo_s = fl_shfl(e_out, src);
        const double ro_s = fl_shfl(rt_out, src);
        const bool ch = fl_elev_warp(e, q_s, root_s, rtp_s, zp_s, eo_s, ro_s);
        if (lane == src) changed |= ch;
    }
#endif
    if (changed) e.flags[FL_FLAG_CHANGED] = 1u;
}

// keys for sorting segment heads by descending nesting height: key = maxh - hgt (heads), FL_NONE otherwise
__global__ void __launch_bounds__(256) k_flow_sort_keys(uint32_t n, const uint32_t* __restrict__ hgt, uint32_t maxh,
                                                         uint32_t* __restrict__ keys) {
    uint32_t q = FL_TID;
    if (q >= n) return;
    const uint32_t h = hgt[q];
    keys[q] = (h == FL_NONE) ? FL_NONE : (maxh - h);
}

// layout rebuild on dynamic segments: exclusive scan input = path length at heads (in index order), 0 elsewhere
__global__ void __launch_bounds__(256) k_head_lengths(uint32_t n, const unsigned long long* __restrict__ pd,
                                                       const uint32_t* __restrict__ plen,
                                                       uint32_t* __restrict__ out) {
    uint32_t q = FL_TID;
    if (q >= n) return;
    out[q] = ((uint32_t)pd[q] == q) ? plen[q] : 0u;
}

__global__ void __launch_bounds__(256) k_newpos_direct(uint32_t n, const unsigned long long* __restrict__ pd,
                                                        const uint32_t* __restrict__ starts,
                                                        uint32_t* __restrict__ newpos) {
    uint32_t q = FL_TID;
    if (q >= n) return;
    const unsigned long long a = pd[q];
    newpos[q] = starts[(uint32_t)a] + (uint32_t)(a >> 32);
}
