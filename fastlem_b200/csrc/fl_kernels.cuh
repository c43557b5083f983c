// fl_kernels.cuh -- device kernels of the terrain solve (sm_100a; also compiled by the FL_EMU host build).
//
// Stage names follow SURVEY.md section 8:  K1 receivers, K2 labels, K3 lake connection, K4 drainage area,
// K5 response time + elevation.  Every floating-point expression below keeps the reference's operand
// order and rounding points (no FMA contraction: the library is compiled with -fmad=false; sqrt and
// division are IEEE-rounded in double precision).
#pragma once
#include "fl_rt.h"

#define FL_NONE 0xFFFFFFFFu
#define FL_KEY_NONE 0xFFFFFFFFFFFFFFFFull

// flag words in Ctx::d_flags
enum { FL_FLAG_LAKE = 0, FL_FLAG_CHANGED = 1, FL_FLAG_JUMP = 2, FL_FLAG_MAXDEPTH = 3, FL_FLAG_REACHED = 4,
       FL_FLAG_PARKED = 5 /* and 6: parked-climb counters of the dataflow sweep */, FL_FLAG_BROKEN = 7,
       FL_FLAG_NCHG = 8 /* receivers changed by K1 */, FL_FLAG_NREGATHER = 9, FL_FLAG_NDIRTY = 10 /* incremental K4 lists */,
       FL_FLAG_TICKET = 11 /* fused K5 launch */, FL_FLAG_K4MAXH = 12 /* largest nesting height met by K4 */,
       FL_FLAG_NROOTS = 13 /* fused incremental K4: dirty tree roots still to finish */,
       FL_N_FLAGS = 192 /* 24..55: level histogram, 64 / 96 / 128: queue counters, 160..167: heads per height below the cut
                           (split K5 sweep, fl_elev.cuh) */ };

#define FL_TID (blockIdx.x * blockDim.x + threadIdx.x)

// ------------------------------------------------------------------------------------------------
// K1: steepest-descent receivers.  reference src/lem/stream_tree.rs:109-137
//   next[i] = i; for non-outlets scan neighbours in adjacency order, strict `e_i > e_j`, strict
//   `slope > steepest` (init 0.0): the first slot attaining the maximum wins.
// Also records the receiver edge length (1.0 for roots: generator.rs:164-171, has_edge(i,i) is false)
// and raises FL_FLAG_LAKE when a non-outlet keeps next[i]==i (stream_tree.rs:155-158).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_receivers(uint32_t n, const uint32_t* __restrict__ row_ptr,
                                                    const uint32_t* __restrict__ col, const double* __restrict__ dist,
                                                    const double* __restrict__ elev,
                                                    const uint8_t* __restrict__ is_outlet, uint32_t* __restrict__ recv,
                                                    double* __restrict__ drecv, uint32_t* __restrict__ flags) {
    uint32_t i = FL_TID;
    if (i >= n) return;
    uint32_t best = i;
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
                }
            }
        }
        if (best == i) atomicOr(&flags[FL_FLAG_LAKE], 1u);
    }
    recv[i] = best;
    drecv[i] = best_d;
}

// ------------------------------------------------------------------------------------------------
// K2: root labels and depths by pointer jumping.  reference src/lem/stream_tree.rs:139-173 computes
// subroot[i] = the root of i in the functional forest `next`; a jump table gives the same labels.
// pd[i] packs (pointer | distance-to-pointer << 32); 64-bit loads/stores keep each pair consistent, so
// the update may run in place and asynchronously.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_jump_init(uint32_t n, const uint32_t* __restrict__ recv,
                                                    unsigned long long* __restrict__ pd) {
    uint32_t i = FL_TID;
    if (i >= n) return;
    uint32_t r = recv[i];
    pd[i] = (unsigned long long)r | ((unsigned long long)(r != i ? 1u : 0u) << 32);
}

__global__ void __launch_bounds__(256) k_jump(uint32_t n, volatile unsigned long long* pd, uint32_t* flags) {
    uint32_t i = FL_TID;
    if (i >= n) return;
    // up to four jumps per launch: every intermediate (pointer, distance) pair another thread may read is consistent,
    // so the rounds need no barrier between them -- fewer launches and read-backs per pointer-jumping loop
    unsigned long long a = pd[i];
    bool moved = false;
#pragma unroll 1
    for (int hop = 0; hop < 4; ++hop) {
        const uint32_t p = (uint32_t)a;
        if (p == i) break;
        const unsigned long long b = pd[p];
        const uint32_t q = (uint32_t)b;
        if (q == p) break;  // p is a root: done
        const uint32_t d = (uint32_t)(a >> 32) + (uint32_t)(b >> 32);
        a = (unsigned long long)q | ((unsigned long long)d << 32);
        moved = true;
    }
    if (moved) {
        pd[i] = a;
        flags[FL_FLAG_JUMP] = 1u;
    }
}

// label = root; depth key = depth if the root is an outlet, FL_NONE otherwise (such nodes are in no
// outlet's basin: generator.rs:149-205 never visits them).
__global__ void __launch_bounds__(256) k_labels_finalize(uint32_t n, const unsigned long long* __restrict__ pd,
                                                          const uint8_t* __restrict__ is_outlet,
                                                          const double* __restrict__ areas,
                                                          uint32_t* __restrict__ label, uint32_t* __restrict__ depth,
                                                          uint32_t* __restrict__ ids, double* __restrict__ A,
                                                          double* __restrict__ rt) {
    uint32_t i = FL_TID;
    if (i >= n) return;
    unsigned long long a = pd[i];
    uint32_t root = (uint32_t)a;
    label[i] = root;
    ids[i] = i;
    if (is_outlet[root]) {
        depth[i] = (uint32_t)(a >> 32);
    } else {
        depth[i] = FL_NONE;
        A[i] = areas[i];  // generator.rs:144-145: untouched initial values
        rt[i] = 0.0;
    }
}

__global__ void __launch_bounds__(256) k_labels_only(uint32_t n, const unsigned long long* __restrict__ pd,
                                                      uint32_t* __restrict__ label) {
    uint32_t i = FL_TID;
    if (i >= n) return;
    label[i] = (uint32_t)pd[i];
}

// ------------------------------------------------------------------------------------------------
// K3: lake connection.  reference src/lem/stream_tree.rs:175-243.
// The sequential flood joins lake basin L through the first pair (popped node i, neighbour j) with
// subroot[j] == L; "first" = (pop order T(i), slot of j in adj(i)).  T is static (fl_flood.cpp), so
// every lake finds its pair with one packed 64-bit atomicMin, then reverses its own in-basin path.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_lake_min(uint32_t n, const uint32_t* __restrict__ row_ptr,
                                                   const uint32_t* __restrict__ col,
                                                   const uint32_t* __restrict__ label,
                                                   const uint8_t* __restrict__ is_outlet,
                                                   const uint32_t* __restrict__ rank,
                                                   unsigned long long* __restrict__ lake_key) {
    uint32_t i = FL_TID;
    if (i >= n) return;
    const uint32_t t = rank[i];
    if (t == FL_NONE) return;
    const uint32_t li = label[i];
    const uint32_t s0 = row_ptr[i], s1 = row_ptr[i + 1];
    for (uint32_t s = s0; s < s1; ++s) {
        const uint32_t l = label[col[s]];
        if (l != li && !is_outlet[l])
            atomicMin(&lake_key[l], ((unsigned long long)t << 32) | (unsigned long long)(s - s0));
    }
}

// has_edge(a,b) of terrain-graph: first match in a's list, else "no edge" (-> 1.0, generator.rs:164-171)
__device__ __forceinline__ double fl_edge_length(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
                                                 const double* __restrict__ dist, uint32_t a, uint32_t b) {
    const uint32_t s1 = row_ptr[a + 1];
    for (uint32_t s = row_ptr[a]; s < s1; ++s)
        if (col[s] == b) return dist[s];
    return 1.0;
}

// one thread per lake root: reverse the path j* -> ... -> root (stream_tree.rs:214-230)
__global__ void __launch_bounds__(256) k_lake_reverse(uint32_t n, const uint32_t* __restrict__ row_ptr,
                                                       const uint32_t* __restrict__ col,
                                                       const double* __restrict__ dist,
                                                       const uint8_t* __restrict__ is_outlet,
                                                       const uint32_t* __restrict__ rank_to_node,
                                                       const unsigned long long* __restrict__ lake_key,
                                                       const uint32_t* __restrict__ label, uint32_t* recv,
                                                       double* drecv) {
    uint32_t l = FL_TID;
    if (l >= n) return;
    if (label[l] != l || is_outlet[l]) return;  // not a lake root
    const unsigned long long key = lake_key[l];
    if (key == FL_KEY_NONE) return;  // no flooded neighbour: the basin is not connected to any outlet
    const uint32_t i = rank_to_node[(uint32_t)(key >> 32)];
    const uint32_t j = col[row_ptr[i] + (uint32_t)key];
    uint32_t k = j, nk = i;
    for (;;) {
        const uint32_t tmp = recv[k];
        recv[k] = nk;
        drecv[k] = fl_edge_length(row_ptr, col, dist, k, nk);
        if (tmp == k) break;
        nk = k;
        k = tmp;
    }
}

// ------------------------------------------------------------------------------------------------
// level bookkeeping (nodes sorted by depth): offs[d] = first sorted position of depth d
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_level_offsets(uint32_t n, const uint32_t* __restrict__ sorted_depth,
                                                        uint32_t* __restrict__ offs, uint32_t* __restrict__ flags) {
    uint32_t p = FL_TID;
    if (p >= n) return;
    const uint32_t d = sorted_depth[p];
    if (d == FL_NONE) return;
    if (p == 0 || sorted_depth[p - 1] != d) offs[d] = p;
    if (p == n - 1 || sorted_depth[p + 1] == FL_NONE) {
        offs[d + 1] = p + 1;
        flags[FL_FLAG_MAXDEPTH] = d;
        flags[FL_FLAG_REACHED] = p + 1;
    }
}

// ------------------------------------------------------------------------------------------------
// K4: drainage area, one tree level per launch (deepest first).  reference generator.rs:154-159 adds
// A[i] into A[next[i]] in reverse BFS order, i.e. node j receives its children in REVERSE adjacency
// order: A[j] = ((a_j + A[c_k]) + ...) + A[c_1].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_area_level(uint32_t begin, uint32_t count, const uint32_t* __restrict__ order,
                                                     const uint32_t* __restrict__ row_ptr,
                                                     const uint32_t* __restrict__ col,
                                                     const uint32_t* __restrict__ recv,
                                                     const double* __restrict__ areas, double* A) {
    uint32_t t = FL_TID;
    if (t >= count) return;
    const uint32_t j = order[begin + t];
    double a = areas[j];
    const uint32_t s0 = row_ptr[j];
    for (uint32_t s = row_ptr[j + 1]; s > s0;) {
        --s;
        const uint32_t c = col[s];
        if (recv[c] == j) a += A[c];
    }
    A[j] = a;
}

// ------------------------------------------------------------------------------------------------
// K5: response time + elevation, one tree level per launch (roots first).  reference generator.rs:162-203
//   celerity = k_i * A_i^0.5 ;  rt_i = 0.0 + (rt_recv + 1.0 / celerity * d_i)
//   z = e_outlet + u_i * max(rt_i - rt_outlet, 0.0);  optional clamp against the receiver's NEW elevation
//   changed |= z != e_i
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_elev_level(uint32_t begin, uint32_t count, int level,
                                                     const uint32_t* __restrict__ order,
                                                     const uint32_t* __restrict__ recv,
                                                     const uint32_t* __restrict__ label,
                                                     const double* __restrict__ drecv, const double* __restrict__ A,
                                                     const double* __restrict__ erod,
                                                     const double* __restrict__ uplift,
                                                     const double* __restrict__ tan_slope, double* elev, double* rt,
                                                     uint32_t* __restrict__ flags) {
    uint32_t t = FL_TID;
    if (t >= count) return;
    const uint32_t i = order[begin + t];
    const uint32_t j = recv[i];
    const double d = drecv[i];  // 1.0 for roots
    const double celerity = erod[i] * sqrt(A[i]);
    const double rt_prev = (level == 0) ? 0.0 : rt[j];  // for a root j == i and rt[i] is still 0.0 (generator.rs:145)
    const double rti = 0.0 + (rt_prev + 1.0 / celerity * d);
    const uint32_t root = label[i];
    const double rt_out = (level == 0) ? rti : rt[root];
    const double e_out = elev[root];  // outlets keep their elevation, so old == new here
    double z = e_out + uplift[i] * fmax(rti - rt_out, 0.0);
    if (tan_slope) {
        const double ms = tan_slope[i];
        if (ms == ms) {  // not NaN: Some(max_slope)
            const double slope = (z - elev[j]) / d;
            if (slope > ms) z = elev[j] + ms * d;
        }
    }
    if (z != elev[i]) flags[FL_FLAG_CHANGED] = 1u;
    elev[i] = z;
    rt[i] = rti;
}
