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
// static per-graph table: rev[s] = slot of i inside adj(col[s]) (first match), 255 if >= 255 / absent;
// also checks that the graph is what terrain-graph's add_edge produces from a triangulation: simple and symmetric
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rev_slots(uint32_t n, const uint32_t* __restrict__ row_ptr,
                                                    const uint32_t* __restrict__ col, const double* __restrict__ dist,
                                                    uint8_t* __restrict__ rev, uint32_t* __restrict__ bad) {
    uint32_t i = FL_TID;
    if (i >= n) return;
    const uint32_t s0 = row_ptr[i], s1 = row_ptr[i + 1];
    uint32_t flaws = 0u;  // 1 self loop, 2 parallel edge, 4 no reverse edge, 8 lengths differ between the directions
    for (uint32_t s = s0; s < s1; ++s) {
        const uint32_t j = col[s];
        if (j == i) flaws |= 1u;
        for (uint32_t t = s0; t < s; ++t)
            if (col[t] == j) flaws |= 2u;
        const uint32_t t0 = row_ptr[j], t1 = row_ptr[j + 1];
        uint32_t r = 255;
        bool found = false;
        for (uint32_t t = t0; t < t1; ++t)
            if (col[t] == i) {
                r = (t - t0) < 255u ? (t - t0) : 255u;
                found = true;
                if (!(dist[t] == dist[s])) flaws |= 8u;
                break;
            }
        if (!found) flaws |= 4u;
        rev[s] = (uint8_t)r;
    }
    if (flaws) atomicOr(bad, flaws);
}

// ------------------------------------------------------------------------------------------------
// K1 (stream_tree.rs:109-137) + child registration: the chosen receiver gets bit `slot of i in its row`
// set in its child mask, so parents can later enumerate children in adjacency order without rescanning
// their neighbours.  Rows longer than 32 ignore the mask and rescan (fl_children_rev).
// ------------------------------------------------------------------------------------------------
// Build-time variants for A/B runs on the device (tools/ab_k1.py): FL_K1_MINBLOCKS n = __launch_bounds__(256, n)
// (0 = ptxas' choice), FL_K1_BATCH = neighbours whose loads are issued together.  Measured on a B200 at 1M sites
// (profiles/r1c_ab_k1_1.txt, _2.txt; results bit-identical): batches of 8 (78 registers, 3 CTAs per SM) 60.5 us per
// launch; 6: 50.3; 4: 51.1; 3 (48 registers, 5 CTAs per SM): 46.6; 2: 50.1; forcing more CTAs per SM through
// __launch_bounds__ only adds spills (batch 8: 60.9 / 69.3 / 70.7 / 82.0 us for 4 / 5 / 6 / 8 CTAs).  With a mean
// degree of 6 a row is two batches of 3.
#ifndef FL_K1_MINBLOCKS
#define FL_K1_MINBLOCKS 0
#endif
#ifndef FL_K1_BATCH
#define FL_K1_BATCH 3
#endif
#if FL_K1_MINBLOCKS > 0 && !defined(FL_EMU)
#define FL_K1_BOUNDS __launch_bounds__(256, FL_K1_MINBLOCKS)
#else
#define FL_K1_BOUNDS __launch_bounds__(256)
#endif
__global__ void FL_K1_BOUNDS k_receivers_mask(uint32_t n, const uint32_t* __restrict__ row_ptr,
                                                         const uint32_t* __restrict__ col,
                                                         const double* __restrict__ dist,
                                                         const uint8_t* __restrict__ rev,
                                                         const double* __restrict__ elev,
                                                         const uint8_t* __restrict__ is_outlet,
                                                         const uint32_t* recv_prev, uint32_t* recv,
                                                         double* __restrict__ drecv,
                                                         uint32_t* cmask, uint32_t* __restrict__ flags,
                                                         uint32_t* __restrict__ chg_node,
                                                         uint32_t* __restrict__ chg_old) {
    uint32_t i = FL_TID;
    if (i >= n) return;
    uint32_t best = i, best_s = FL_NONE;
    double best_d = 1.0;
    if (!is_outlet[i]) {
        const double ei = elev[i];
        double steepest = 0.0;
        const uint32_t s0 = row_ptr[i], s1 = row_ptr[i + 1];
        // FL_K1_BATCH neighbours at a time: their ids and edge lengths first, then the elevation gathers -- all loads of
        // a batch are in flight together; the comparisons then run in adjacency order (first slot wins ties)
        for (uint32_t sb = s0; sb < s1; sb += (uint32_t)FL_K1_BATCH) {
            uint32_t j[FL_K1_BATCH];
            double d[FL_K1_BATCH], ej[FL_K1_BATCH];
#pragma unroll
            for (int k = 0; k < FL_K1_BATCH; ++k) {
                const uint32_t s = sb + (uint32_t)k;
                const bool ok = s < s1;
                j[k] = ok ? col[s] : i;  // padding: the site itself (never lower than itself)
                d[k] = ok ? dist[s] : 1.0;
            }
#pragma unroll
            for (int k = 0; k < FL_K1_BATCH; ++k) ej[k] = elev[j[k]];
#pragma unroll
            for (int k = 0; k < FL_K1_BATCH; ++k) {
                if (ei > ej[k]) {
                    const double slope = (ei - ej[k]) / d[k];
                    if (slope > steepest) {
                        steepest = slope;
                        best = j[k];
                        best_d = d[k];
                        best_s = sb + (uint32_t)k;
                    }
                }
            }
        }
        if (best == i) atomicOr(&flags[FL_FLAG_LAKE], 1u);
    }
    if (chg_node) {  // incremental K4: sites whose receiver differs from the previous iteration's (final) one
        const uint32_t old = recv_prev[i];  // (recv_prev may be recv itself)
        if (old != best) {
            const uint32_t k = atomicAdd(&flags[FL_FLAG_NCHG], 1u);
            chg_node[k] = i;
            chg_old[k] = old;
        }
    }
    recv[i] = best;
    drecv[i] = best_d;
    if (best_s != FL_NONE) {
        const uint32_t r = rev[best_s];
        if (r < 32u) atomicOr(&cmask[best], 1u << r);
    }
}

#ifndef FL_EMU
// ------------------------------------------------------------------------------------------------
// K1 with the CSR stream staged through shared memory by the bulk-copy engine (cp.async.bulk + mbarrier; SASS: UBLKCP).
//
// The thread-per-row kernel above is bound by dependent round trips, not by bytes: row_ptr -> col / dist of a batch ->
// the elevation gathers -> the next batch (5-6 latencies in a row per thread, 1280 threads per SM).  Here a persistent
// CTA walks tiles of FL_K1B_ROWS consecutive rows; the tile's contiguous col / dist span (rows are contiguous in CSR) is
// copied global -> shared by ONE bulk copy each, issued one tile ahead into the other stage, so when a tile starts its
// neighbour ids are a shared-memory read away and a thread issues ALL its elevation gathers at once: one exposed
// round trip per row.  Same comparisons, same order (stream_tree.rs:109-137): results are bit-identical.
// A span longer than a stage (FL_K1B_CAP slots; mean degree 6 -> 1536) is read from global memory directly.
// ------------------------------------------------------------------------------------------------
#define FL_K1B_ROWS 256u
#define FL_K1B_CAP 2048u
#define FL_K1B_GATHER 8
// one stage: dist | col | rev of the tile's span, then the tile's own rows: row_ptr (257 used), elev, is_outlet, recv_prev
#define FL_K1B_OFF_COL (FL_K1B_CAP * 8u)
#define FL_K1B_OFF_REV (FL_K1B_OFF_COL + FL_K1B_CAP * 4u)
#define FL_K1B_OFF_ROWP (FL_K1B_OFF_REV + FL_K1B_CAP)
#define FL_K1B_OFF_ELEV (FL_K1B_OFF_ROWP + 272u * 4u)
#define FL_K1B_OFF_OUTLET (FL_K1B_OFF_ELEV + FL_K1B_ROWS * 8u)
#define FL_K1B_OFF_PREV (FL_K1B_OFF_OUTLET + FL_K1B_ROWS)
#define FL_K1B_STAGE_BYTES (FL_K1B_OFF_PREV + FL_K1B_ROWS * 4u)
#define FL_K1B_SMEM (2u * FL_K1B_STAGE_BYTES + 64u)

__device__ __forceinline__ uint32_t fl_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fl_mbar_init(void* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fl_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fl_mbar_expect_tx(void* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fl_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fl_mbar_wait(void* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(fl_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fl_bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(fl_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(fl_smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(256) k_receivers_bulk(uint32_t n, uint32_t n_tiles, const uint32_t* __restrict__ row_ptr,
                                                         const uint32_t* __restrict__ col,
                                                         const double* __restrict__ dist,
                                                         const uint8_t* __restrict__ rev,
                                                         const double* __restrict__ elev,
                                                         const uint8_t* __restrict__ is_outlet,
                                                         const uint32_t* recv_prev, uint32_t* recv,
                                                         double* __restrict__ drecv,
                                                         uint32_t* cmask, uint32_t* __restrict__ flags,
                                                         uint32_t* __restrict__ chg_node,
                                                         uint32_t* __restrict__ chg_old) {
    extern __shared__ __align__(128) unsigned char k1b_smem[];
    unsigned long long* bars = (unsigned long long*)(k1b_smem + 2u * FL_K1B_STAGE_BYTES);
    uint32_t* meta = (uint32_t*)(k1b_smem + 2u * FL_K1B_STAGE_BYTES + 16u);  // per stage: span start, span staged?
    const uint32_t tid = threadIdx.x;
    if (tid == 0u) {
        fl_mbar_init(&bars[0], 1u);
        fl_mbar_init(&bars[1], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // thread 0: the span of the tile to be copied next (loaded one iteration before it is needed)
    uint32_t nb = 0u, ne = 0u;
    auto bounds = [&](uint32_t tile) {
        const unsigned long long t0 = (unsigned long long)tile * FL_K1B_ROWS;
        const unsigned long long t1 = t0 + FL_K1B_ROWS < n ? t0 + FL_K1B_ROWS : n;
        nb = row_ptr[t0];
        ne = row_ptr[t1];
    };
    // everything a tile reads except the neighbours' elevations, by the copy engine: its own rows (row_ptr, elev,
    // is_outlet, previous receivers) and its CSR span [nb, ne), the span's start aligned down to 16 slots (16 bytes in
    // every array).  Sizes are multiples of 16 bytes; what they read beyond the tile lies inside the same allocation.
    auto issue = [&](uint32_t tile, uint32_t stage) {
        const unsigned long long t0 = (unsigned long long)tile * FL_K1B_ROWS;
        const uint32_t rows = (uint32_t)(t0 + FL_K1B_ROWS < n ? FL_K1B_ROWS : n - t0);
        const uint32_t start = nb & ~15u;
        const uint32_t cnt = (ne - start + 15u) & ~15u;
        const bool ok = cnt > 0u && cnt <= FL_K1B_CAP;
        meta[2u * stage] = start;
        meta[2u * stage + 1u] = ok ? 1u : 0u;
        unsigned char* base = k1b_smem + stage * FL_K1B_STAGE_BYTES;
        const uint32_t b_rowp = ((rows + 1u) * 4u + 15u) & ~15u, b_elev = (rows * 8u + 15u) & ~15u,
                       b_out = (rows + 15u) & ~15u, b_prev = (rows * 4u + 15u) & ~15u;
        fl_mbar_expect_tx(&bars[stage], b_rowp + b_elev + b_out + (chg_node ? b_prev : 0u) + (ok ? cnt * 13u : 0u));
        fl_bulk_g2s(base + FL_K1B_OFF_ROWP, row_ptr + t0, b_rowp, &bars[stage]);
        fl_bulk_g2s(base + FL_K1B_OFF_ELEV, elev + t0, b_elev, &bars[stage]);
        fl_bulk_g2s(base + FL_K1B_OFF_OUTLET, is_outlet + t0, b_out, &bars[stage]);
        if (chg_node) fl_bulk_g2s(base + FL_K1B_OFF_PREV, recv_prev + t0, b_prev, &bars[stage]);
        if (ok) {
            fl_bulk_g2s(base, dist + start, cnt * 8u, &bars[stage]);
            fl_bulk_g2s(base + FL_K1B_OFF_COL, col + start, cnt * 4u, &bars[stage]);
            fl_bulk_g2s(base + FL_K1B_OFF_REV, rev + start, cnt, &bars[stage]);
        }
    };
    uint32_t tile = blockIdx.x;
    if (tid == 0u && tile < n_tiles) {
        bounds(tile);
        issue(tile, 0u);
        if (tile + gridDim.x < n_tiles) bounds(tile + gridDim.x);
    }
    uint32_t parity0 = 0u, parity1 = 0u, stage = 0u;
    bool lake = false;
    for (; tile < n_tiles; tile += gridDim.x, stage ^= 1u) {
        if (tid == 0u && tile + gridDim.x < n_tiles) {  // (the other stage was released by the barrier that ended the previous tile)
            issue(tile + gridDim.x, stage ^ 1u);
            const unsigned long long after = (unsigned long long)tile + 2ull * gridDim.x;
            if (after < n_tiles) bounds((uint32_t)after);
        }
        const unsigned long long i64 = (unsigned long long)tile * FL_K1B_ROWS + tid;
        const bool active = i64 < n;
        const uint32_t i = (uint32_t)i64;
        fl_mbar_wait(&bars[stage], stage ? parity1 : parity0);
        if (stage) parity1 ^= 1u; else parity0 ^= 1u;
        const unsigned char* base = k1b_smem + stage * FL_K1B_STAGE_BYTES;
        const uint32_t start = meta[2u * stage];
        const bool staged = meta[2u * stage + 1u] != 0u;
        uint32_t best = i, best_s = FL_NONE, best_r = 0xFFu;
        double best_d = 1.0;
        if (active) {
            const uint32_t s0 = ((const uint32_t*)(base + FL_K1B_OFF_ROWP))[tid];
            const uint32_t s1 = ((const uint32_t*)(base + FL_K1B_OFF_ROWP))[tid + 1u];
            const bool outlet = (base + FL_K1B_OFF_OUTLET)[tid] != 0;
            const double ei = ((const double*)(base + FL_K1B_OFF_ELEV))[tid];
            if (!outlet) {
                const double* sd = (const double*)base + (s0 - start);
                const uint32_t* sc = (const uint32_t*)(base + FL_K1B_OFF_COL) + (s0 - start);
                const uint32_t deg = s1 - s0;
                double steepest = 0.0;
                for (uint32_t sb = 0u; sb < deg; sb += (uint32_t)FL_K1B_GATHER) {
                    uint32_t j[FL_K1B_GATHER];
                    double ej[FL_K1B_GATHER];
#pragma unroll
                    for (int k = 0; k < FL_K1B_GATHER; ++k) {
                        const uint32_t o = sb + (uint32_t)k;
                        j[k] = o < deg ? (staged ? sc[o] : col[s0 + o]) : i;
                    }
#pragma unroll
                    for (int k = 0; k < FL_K1B_GATHER; ++k)  // all gathers of the row in flight together
                        ej[k] = (sb + (uint32_t)k < deg) ? elev[j[k]] : ei;  // padding: never lower than the site itself
#pragma unroll
                    for (int k = 0; k < FL_K1B_GATHER; ++k) {
                        if (ei > ej[k]) {
                            const uint32_t o = sb + (uint32_t)k;
                            const double d = staged ? sd[o] : dist[s0 + o];
                            const double slope = (ei - ej[k]) / d;
                            if (slope > steepest) {
                                steepest = slope;
                                best = j[k];
                                best_d = d;
                                best_s = s0 + o;
                            }
                        }
                    }
                }
                if (best == i) lake = true;
                else best_r = staged ? (base + FL_K1B_OFF_REV)[best_s - start] : rev[best_s];
            }
            if (chg_node) {  // incremental K4: sites whose receiver differs from the previous iteration's
                const uint32_t old = ((const uint32_t*)(base + FL_K1B_OFF_PREV))[tid];
                if (old != best) {
                    const uint32_t k = atomicAdd(&flags[FL_FLAG_NCHG], 1u);
                    chg_node[k] = i;
                    chg_old[k] = old;
                }
            }
            recv[i] = best;
            drecv[i] = best_d;
            if (best_r < 32u) atomicOr(&cmask[best], 1u << best_r);
        }
        __syncthreads();  // every read of this stage is done: it may be refilled
    }
    if (lake) atomicOr(&flags[FL_FLAG_LAKE], 1u);
}
#endif

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
    const uint32_t *recv, *cmask, *rank, *orig_of, *lvl;
    uint32_t *recv_n, *cmask_n, *rank_n, *orig_of_n, *lvl_n;
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
    if (a.lvl) a.lvl_n[p] = a.lvl[q];
    a.is_outlet_n[p] = a.is_outlet[q];
}

__global__ void __launch_bounds__(256) k_rank_inverse(uint32_t n, const uint32_t* __restrict__ rank,
                                                       uint32_t* __restrict__ rank_to_node) {
    uint32_t q = FL_TID;
    if (q >= n) return;
    const uint32_t t = rank[q];
    if (t != FL_NONE) rank_to_node[t] = q;
}

// a[k] <- newpos[a[k]]  (a list of site ids follows a renumbering)
__global__ void __launch_bounds__(256) k_map_list(uint32_t count, const uint32_t* __restrict__ newpos,
                                                   uint32_t* __restrict__ a) {
    uint32_t k = FL_TID;
    if (k < count) a[k] = newpos[a[k]];
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

