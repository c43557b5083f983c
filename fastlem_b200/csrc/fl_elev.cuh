// fl_elev.cuh -- K5, response time + elevation + max_slope clamp + `changed` (generator.rs:162-203), split by the
// nesting height of the segments (a segment = a run of positions q, q+1, ... with recv[q+1] == q, fl_flow.cuh; its
// nesting height hgt = longest chain of segment-to-segment hand-offs below it, left at every segment head by K4).
//
// The sweep runs outlet -> upstream; a site needs nothing but its receiver's NEW values.  The forest has two very
// different parts: a sparse TOP (few, long segments: trunks and their tributaries; a long chain of dependent
// hand-offs) and a populous BOTTOM (most segments are one to three sites long and have no or only leaf children).
//
//   k_elev_plan  one pass over hgt[]: level histogram (the host picks the cut height from it for the next iteration);
//                every head with hgt >= cut registers itself in its receiver's push mask `pmask` (bit = its slot in the
//                receiver's row); roots with hgt >= cut are computed right here and seed the queue.
//   k_elev_top   segments with hgt >= cut, ONE persistent launch, dataflow: a queue of segment starts whose entries
//                carry the receiver's new values (root, rt, z), so a consumer needs no gather from the producer's
//                arrays.  One warp per entry: 32-site windows (coalesced loads, FL_PDEPTH windows prefetched, the
//                serial additions staged through shared memory and run identically by every lane); the children
//                registered in pmask are pushed window by window, i.e. a tributary starts as soon as the site it
//                joins is written, not when the whole trunk is done.  Queue order is a topological order (an entry is
//                appended by the segment that holds its receiver), tickets are handed out in queue order to resident
//                warps, so a waiting warp only ever waits for warps that are running.
//   k_elev_low   heights cut-1 ... 0, one launch per height over ALL positions (a thread whose site heads a segment of
//                that height walks it; the warp finishes the rare long ones together): no sorting, no lists, loads stay
//                in position order; the receiver's values are complete because its segment is strictly higher.
//
// Every floating-point operation is the reference's, in the reference's order:
//     celerity = k_i * A_i^0.5 ;  rt_i = 0.0 + (rt_recv + 1.0 / celerity * d_i)                 generator.rs:162-174
//     z = e_outlet + u_i * max(rt_i - rt_outlet, 0.0) ; clamp against the receiver's NEW elevation   :177-203
// Trees whose root is not an outlet are never visited (generator.rs:149): their sites get root_of = FL_NONE and keep
// their elevation.
//
// Ordering between threads of k_elev_top: producer = payload stores, st.release.gpu of the entry's flag word;
// consumer = ld.acquire.gpu of the flag word, then the payload loads (PTX memory model, release / acquire pattern).
#pragma once
#include "fl_flow.cuh"

#define FL_PB 4       // sites per thread-level batch (loads issued together)
#define FL_PSHORT 8   // k_elev_low: sites a thread walks alone before the warp finishes the segment together
#ifndef FL_PDEPTH
#define FL_PDEPTH 3   // windows of a segment kept in flight
#endif
#define FL_CUT_MAX 8u  // largest cut height (= most k_elev_low launches per iteration)
#ifndef FL_TOP_WAIT_LIMIT_NS
// how long a warp of k_elev_top polls for its queue entry before the sweep is declared broken.  A guard against hanging,
// not a schedule: idle warps wait as long as the kernel runs, so the limit bounds the kernel's run time.  It used to be
// a poll count (2^24, a few seconds) and fired in a two-process run with four contexts per GPU on 4M-site models
// (profiles/r2v_bench_2gpu_failed.err) although the same members pass in one process; wall clock, and generous, now.
#define FL_TOP_WAIT_LIMIT_NS 120000000000ull
#endif
#ifndef FL_TOP_EARLY_WINDOW
#define FL_TOP_EARLY_WINDOW 1  // build-time A/B switch: k_elev_top requests a run's first window with the entry's payload
#endif

// words of d_flags
// (the three queue counters sit on cache lines of their own: producers hit tail, consumers head, finishers done)
enum { FLQ_HIST = 24 /* 32 bins: heads per nesting height (31 = 31+) */, FLQ_TAIL = 64, FLQ_HEAD = 96, FLQ_DONE = 128,
       FLQ_LOW = 160 /* FL_CUT_MAX words: heads listed per height below the cut */ };

struct FlQEntry {  // 32 bytes = one sector
    unsigned long long flag;  // (epoch << 32) | first site of the run; written last (release)
    uint32_t root;            // tree root (an outlet), FL_NONE = tree without outlet
    uint32_t pad;
    double rt_p;              // response time of the run's receiver
    double z_p;               // NEW elevation of the run's receiver
};

struct FlSplit {
    uint32_t n;
    const uint32_t* row_ptr;
    const uint32_t* col;
    const uint8_t* rev;
    const uint32_t* recv;
    const uint32_t* cmask;
    const uint8_t* is_outlet;
    const uint32_t* hgt;
    const double* drecv;
    const double* erod;
    const double* A;
    const double* uplift;
    const double* tan_slope;  // may be null
    double* tcel;      // 1.0 / (k * A^0.5) * d per site: written by k_elev_plan, read by the sweeps
    double* elev;
    double* rt;
    uint32_t* root_of;
    uint32_t* pmask;   // per site: children that start a queue entry (consumed and cleared by k_elev_top)
    FlQEntry* queue;   // n entries
    uint32_t epoch;    // number of this sweep: entries of earlier sweeps never match, the queue is never cleared
    uint32_t cut;      // segments with hgt >= cut go through the queue
    uint32_t* flags;
    // the heads below the cut as (head, receiver) pairs, one list per height, filled by k_elev_plan (flags[FLQ_LOW + l] =
    // length).  A segment of height l has a chain of l segments below it and segments are disjoint, so there are at
    // most n / (l + 1) heads of height l: the lists lie behind each other in low_list at fl_low_region(n, l).
    uint2* low_list;
};
#ifndef FL_LOW_CHUNK
#define FL_LOW_CHUNK 512u  // positions per block of k_elev_plan (build-time A/B switch; 512 measured best at 1M sites, no difference at 16M: profiles/r2n_ab.txt)
#endif
__host__ __device__ inline size_t fl_low_region(uint32_t n, uint32_t level) {
    size_t at = 0;
    for (uint32_t l = 0; l < level; ++l) at += (size_t)n / (l + 1u) + 1u;
    return at;
}

__device__ __forceinline__ unsigned long long flq_flag(uint32_t epoch, uint32_t site) {
    return ((unsigned long long)epoch << 32) | (unsigned long long)site;
}
#ifdef FL_EMU
__device__ __forceinline__ unsigned long long flq_ld_acquire(const unsigned long long* p) { return *p; }
__device__ __forceinline__ void flq_st_release(unsigned long long* p, unsigned long long v) { *p = v; }
#else
__device__ __forceinline__ unsigned long long flq_ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void flq_st_release(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#endif

// append a run start; `release`: other threads of the SAME launch consume it (k_elev_top), else the next launch does
__device__ __forceinline__ void flq_put(const FlSplit& e, uint32_t slot, uint32_t site, uint32_t root, double rt_p,
                                        double z_p, bool release) {
    FlQEntry* q = &e.queue[slot];
    q->root = root;
    q->rt_p = rt_p;
    q->z_p = z_p;
    if (release) flq_st_release(&q->flag, flq_flag(e.epoch, site));
    else q->flag = flq_flag(e.epoch, site);
}

// what a run starts from
struct FlSegStart {
    uint32_t root;
    double rt_prev, z_prev, e_out, rt_out;
};

// generator.rs:172-173 for one site
__device__ __forceinline__ double fl_celerity_term(const FlSplit& e, uint32_t i, double d) {
    const double celerity = e.erod[i] * sqrt(e.A[i]);
    return 1.0 / celerity * d;
}

// up to B sites of one run by one thread, starting at q (q is NOT a tree root: roots are handled by fl_root_site);
// returns true when the run ended inside the batch.  A dead run (root == FL_NONE) only marks its sites.
template <int B>
struct FlBatch {
    double d[B], t[B], up[B], eo[B], ms[B];
    uint32_t nx[B];
    uint32_t nb;
};
// the loads of a batch: they depend on nothing but the position, so a caller may issue them before it knows what the
// run starts from (`live` = false skips the values of a dead run when that is already known)
template <int B>
__device__ __forceinline__ void fl_batch_load(const FlSplit& e, uint32_t q, bool live, FlBatch<B>& b) {
    b.nb = e.n - q < (uint32_t)B ? e.n - q : (uint32_t)B;
#pragma unroll
    for (int k = 0; k < B; ++k) {
        b.d[k] = 1.0; b.t[k] = 0.0; b.up[k] = 0.0; b.eo[k] = 0.0; b.ms[k] = 0.0; b.nx[k] = FL_NONE;
        if ((uint32_t)k < b.nb) {
            const uint32_t i = q + k;
            b.nx[k] = (i + 1u < e.n) ? e.recv[i + 1u] : FL_NONE;
            if (live) {
                b.t[k] = e.tcel[i];
                b.up[k] = e.uplift[i];
                b.eo[k] = e.elev[i];
                if (e.tan_slope) { b.ms[k] = e.tan_slope[i]; b.d[k] = e.drecv[i]; }
            }
        }
    }
}
template <int B>
__device__ __forceinline__ bool fl_batch_compute(const FlSplit& e, uint32_t& q, FlSegStart& s, bool& changed,
                                                 const FlBatch<B>& b) {
    const bool live = s.root != FL_NONE;
    bool ended = false;
    uint32_t walked = 0u;
#pragma unroll
    for (int k = 0; k < B; ++k) {
        if (!ended && (uint32_t)k < b.nb) {
            const uint32_t i = q + k;
            ++walked;
            if (live) {
                const double rti = 0.0 + (s.rt_prev + b.t[k]);
                double z = s.e_out + b.up[k] * fmax(rti - s.rt_out, 0.0);
                if (e.tan_slope) {
                    if (b.ms[k] == b.ms[k]) {  // not NaN: Some(max_slope)
                        const double slope = (z - s.z_prev) / b.d[k];
                        if (slope > b.ms[k]) z = s.z_prev + b.ms[k] * b.d[k];
                    }
                }
                changed |= (z != b.eo[k]);
                e.elev[i] = z;
                e.rt[i] = rti;
                s.rt_prev = rti;
                s.z_prev = z;
            }
            e.root_of[i] = s.root;
            if (b.nx[k] != i) ended = true;
        }
    }
    q += walked;  // one past the last site written
    return ended || q >= e.n;
}
// up to B sites of one run by one thread, starting at q (q is NOT a tree root: roots are handled by fl_root_site);
// returns true when the run ended inside the batch.  A dead run (root == FL_NONE) only marks its sites.
template <int B>
__device__ __forceinline__ bool fl_run_batch(const FlSplit& e, uint32_t& q, FlSegStart& s, bool& changed) {
    FlBatch<B> b;
    fl_batch_load<B>(e, q, s.root != FL_NONE, b);
    return fl_batch_compute<B>(e, q, s, changed, b);
}

// A tree root h (recv[h] == h).  An outlet: rt = 0.0 + (0.0 + term), the elevation formula with e_outlet = its own
// old elevation and rt_outlet = its own new rt, the clamp against its own old elevation (has_edge(i,i) is false);
// sites after it read elevations[outlet] AFTER this update.  Not an outlet: the tree is never visited.
__device__ __forceinline__ FlSegStart fl_root_site(const FlSplit& e, uint32_t h, double t, bool& changed) {
    FlSegStart s;
    s.root = e.is_outlet[h] ? h : FL_NONE;
    s.rt_prev = 0.0; s.z_prev = 0.0; s.e_out = 0.0; s.rt_out = 0.0;
    if (s.root == FL_NONE) { e.root_of[h] = FL_NONE; return s; }
    const double eold = e.elev[h];
    const double rti = 0.0 + (0.0 + t);
    double z = eold + e.uplift[h] * fmax(rti - rti, 0.0);
    if (e.tan_slope) {
        const double ms = e.tan_slope[h];
        if (ms == ms) {
            const double d = e.drecv[h];
            const double slope = (z - eold) / d;
            if (slope > ms) z = eold + ms * d;
        }
    }
    changed |= (z != eold);
    e.elev[h] = z;
    e.rt[h] = rti;
    e.root_of[h] = h;
    s.rt_prev = rti; s.z_prev = z; s.e_out = z; s.rt_out = rti;
    return s;
}

// ------------------------------------------------------------------------------------------------
// plan: histogram, push masks, the roots of the top trees
// ------------------------------------------------------------------------------------------------
// a head q of height h >= cut (receiver p, celerity term tq) in the plan pass: a root is computed here and seeds the
// queue with the run behind it and its top children; any other head registers in its receiver's push mask
__device__ __forceinline__ void fl_plan_top_head(const FlSplit& e, uint32_t q, uint32_t p, double tq, bool& changed) {
    if (p == q) {
        // a top tree: the root site itself, then the run behind it and its top children as queue entries
        const FlSegStart s = fl_root_site(e, q, tq, changed);
        const bool chain = (q + 1u < e.n) && (e.recv[q + 1u] == q);
        if (chain) flq_put(e, atomicAdd(&e.flags[FLQ_TAIL], 1u), q + 1u, s.root, s.rt_prev, s.z_prev, false);
        const uint32_t s0 = e.row_ptr[q];
        uint32_t m = e.cmask[q];
        while (m) {
            const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
            m &= m - 1u;
            const uint32_t c = e.col[s0 + b];
            if (chain && c == q + 1u) continue;
            const uint32_t hc = e.hgt[c];
            if (hc != FL_NONE && hc >= e.cut)
                flq_put(e, atomicAdd(&e.flags[FLQ_TAIL], 1u), c, s.root, s.rt_prev, s.z_prev, false);
        }
    } else if (e.recv[p] != p) {  // (a root pushes its own top children, above)
        uint32_t bit = 32u;
        for (uint32_t sl = e.row_ptr[q]; sl < e.row_ptr[q + 1u]; ++sl)
            if (e.col[sl] == p) { bit = e.rev[sl]; break; }
        if (bit < 32u) atomicOr(&e.pmask[p], 1u << bit);
        else atomicOr(&e.flags[FL_FLAG_BROKEN], 16u);
    }
}

#ifdef FL_EMU
__global__ void __launch_bounds__(256) k_elev_plan(FlSplit e) {
    const uint32_t q = FL_TID;
    if (q >= e.n) return;
    bool changed = false;
    const double tq = fl_celerity_term(e, q, e.drecv[q]);  // (fully parallel: the division and the square root stay
    e.tcel[q] = tq;                                        //  out of the serial chains)
    const uint32_t h = e.hgt[q];
    if (h != FL_NONE) {  // q heads a segment
        atomicAdd(&e.flags[FLQ_HIST + (h < 31u ? h : 31u)], 1u);
        if (h >= e.cut) fl_plan_top_head(e, q, e.recv[q], tq, changed);
    }
    if (changed) e.flags[FL_FLAG_CHANGED] = 1u;
}
#else
// One block per chunk of FL_LOW_CHUNK positions, FL_PLAN_PER_THREAD per thread.  All loads of a thread's sites go out
// first; then the celerity terms (the division and the square root stay out of the serial chains), the rare heads at or
// above the cut, and finally the heads below the cut are appended to the per-height lists: per warp one ballot per
// height (position order inside a warp's 32 positions), per block ONE reservation per height in the global list -- so
// the per-height launches of k_elev_low neither scan the heights again nor run half-empty warps.
#define FL_PLAN_PER_THREAD (FL_LOW_CHUNK / 256u)
__global__ void __launch_bounds__(256) k_elev_plan(FlSplit e) {
    __shared__ uint32_t hist[32];
    __shared__ uint32_t lcount[FL_CUT_MAX];
    __shared__ unsigned long long gbase[FL_CUT_MAX];
    if (threadIdx.x < 32u) hist[threadIdx.x] = 0u;
    if (threadIdx.x < FL_CUT_MAX) lcount[threadIdx.x] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned long long base = (unsigned long long)blockIdx.x * FL_LOW_CHUNK;
    bool changed = false;
    double dq[FL_PLAN_PER_THREAD], kq[FL_PLAN_PER_THREAD], aq[FL_PLAN_PER_THREAD];
    uint32_t hq[FL_PLAN_PER_THREAD], rcv[FL_PLAN_PER_THREAD];
#pragma unroll
    for (uint32_t j = 0; j < FL_PLAN_PER_THREAD; ++j) {
        const unsigned long long q64 = base + j * 256u + threadIdx.x;
        dq[j] = 1.0; kq[j] = 1.0; aq[j] = 1.0; hq[j] = FL_NONE; rcv[j] = 0u;
        if (q64 < e.n) {
            const uint32_t q = (uint32_t)q64;
            dq[j] = e.drecv[q]; kq[j] = e.erod[q]; aq[j] = e.A[q]; hq[j] = e.hgt[q]; rcv[j] = e.recv[q];
        }
    }
#pragma unroll
    for (uint32_t j = 0; j < FL_PLAN_PER_THREAD; ++j) {
        const unsigned long long q64 = base + j * 256u + threadIdx.x;
        if (q64 < e.n) {
            const double celerity = kq[j] * sqrt(aq[j]);  // generator.rs:172-173
            const double tq = 1.0 / celerity * dq[j];
            e.tcel[(uint32_t)q64] = tq;
            if (hq[j] != FL_NONE && hq[j] >= e.cut) {
                atomicAdd(&hist[hq[j] < 31u ? hq[j] : 31u], 1u);
                fl_plan_top_head(e, (uint32_t)q64, rcv[j], tq, changed);
            }
        }
    }
    uint32_t slot[FL_PLAN_PER_THREAD];  // rank among the block's heads of the same height (below the cut)
#pragma unroll
    for (uint32_t j = 0; j < FL_PLAN_PER_THREAD; ++j) {
        slot[j] = 0u;
        for (uint32_t l = 0; l < e.cut; ++l) {
            const uint32_t peers = __ballot_sync(FL_FULL, hq[j] == l);
            if (!peers) continue;
            const int leader = __ffs((int)peers) - 1;
            uint32_t first = 0u;
            if (lane == leader) first = atomicAdd(&lcount[l], (uint32_t)__popc(peers));
            first = __shfl_sync(FL_FULL, first, leader);
            if (hq[j] == l) slot[j] = first + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        }
    }
    __syncthreads();
    // one reservation per block and height in the global lists
    if (threadIdx.x < FL_CUT_MAX) {
        const uint32_t l = threadIdx.x, cnt = lcount[l];
        const uint32_t first = cnt ? atomicAdd(&e.flags[FLQ_LOW + l], cnt) : 0u;
        gbase[l] = (unsigned long long)fl_low_region(e.n, l) + first;
        // cannot happen: more heads of a height than sites allow (the list would run into the next height's)
        if ((unsigned long long)first + cnt > (unsigned long long)e.n / (l + 1u) + 1ull) { gbase[l] = ~0ull; atomicOr(&e.flags[FL_FLAG_BROKEN], 64u); }
        if (cnt) atomicAdd(&e.flags[FLQ_HIST + l], cnt);
    }
    __syncthreads();
#pragma unroll
    for (uint32_t j = 0; j < FL_PLAN_PER_THREAD; ++j)
        if (hq[j] < e.cut) {  // (FL_NONE is not)
            const unsigned long long g = gbase[hq[j]];
            if (g != ~0ull) e.low_list[g + slot[j]] = make_uint2((uint32_t)(base + j * 256u + threadIdx.x), rcv[j]);
        }
    if (changed) e.flags[FL_FLAG_CHANGED] = 1u;
    if (threadIdx.x < 32u && hist[threadIdx.x]) atomicAdd(&e.flags[FLQ_HIST + threadIdx.x], hist[threadIdx.x]);
}
#endif

// children of site i registered in its push mask -> queue entries carrying i's new values; clears the mask
template <class F>
__device__ __forceinline__ void fl_push_masked(const FlSplit& e, uint32_t i, uint32_t s0, uint32_t m, F&& put) {
    while (m) {
        const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
        m &= m - 1u;
        put(e.col[s0 + b]);
    }
    e.pmask[i] = 0u;
}

#ifdef FL_EMU
// host emulation: one thread drains the queue in FIFO order (receivers before their children)
__global__ void k_elev_top(FlSplit e) {
    if (FL_TID != 0u) return;
    uint32_t head = 0, tail = e.flags[FLQ_TAIL];
    bool changed = false;
    while (head < tail) {
        const FlQEntry ent = e.queue[head++];
        if ((uint32_t)(ent.flag >> 32) != e.epoch) { atomicOr(&e.flags[FL_FLAG_BROKEN], 2u); return; }
        uint32_t q = (uint32_t)ent.flag;
        FlSegStart s;
        s.root = ent.root; s.rt_prev = ent.rt_p; s.z_prev = ent.z_p;
        s.e_out = s.root != FL_NONE ? e.elev[s.root] : 0.0;
        s.rt_out = s.root != FL_NONE ? e.rt[s.root] : 0.0;
        for (bool ended = false; !ended;) {
            const uint32_t first = q;
            ended = fl_run_batch<FL_PB>(e, q, s, changed);
            for (uint32_t i = first; i < q; ++i) {
                const uint32_t m = e.pmask[i];
                if (!m) continue;
                const double rt_i = s.root != FL_NONE ? e.rt[i] : 0.0, z_i = s.root != FL_NONE ? e.elev[i] : 0.0;
                fl_push_masked(e, i, e.row_ptr[i], m, [&](uint32_t c) {
                    if (tail >= e.n) { atomicOr(&e.flags[FL_FLAG_BROKEN], 2u); return; }
                    flq_put(e, tail++, c, s.root, rt_i, z_i, false);
                });
            }
        }
    }
    e.flags[FLQ_TAIL] = tail;
    e.flags[FLQ_HEAD] = head;
    e.flags[FLQ_DONE] = tail;
    if (changed) e.flags[FL_FLAG_CHANGED] = 1u;
}
#else

// ------------------------------------------------------------------------------------------------
// 32-site windows of a run: lane l holds site base + l
// ------------------------------------------------------------------------------------------------
struct FlPWin {
    double t, up, eold, ms, d;
    uint32_t nx, pm, s0;
    bool valid;
};

template <bool PUSH>
__device__ __forceinline__ FlPWin fl_pwin_load(const FlSplit& e, uint32_t base, int lane, bool live) {
    FlPWin w;
    w.t = 0.0; w.up = 0.0; w.eold = 0.0; w.ms = 0.0; w.d = 1.0; w.nx = FL_NONE; w.pm = 0u; w.s0 = 0u;
    const unsigned long long i64 = (unsigned long long)base + (unsigned)lane;
    w.valid = i64 < e.n;
    if (w.valid) {
        const uint32_t i = (uint32_t)i64;
        w.nx = (i + 1u < e.n) ? e.recv[i + 1u] : FL_NONE;
        if (PUSH) { w.pm = e.pmask[i]; w.s0 = e.row_ptr[i]; }
        if (live) {
            w.t = e.tcel[i];
            w.up = e.uplift[i];
            w.eold = e.elev[i];
            if (e.tan_slope) { w.ms = e.tan_slope[i]; w.d = e.drecv[i]; }
        }
    }
    return w;
}

// sites of the window that belong to the run (it ends at the first site whose successor is not chained to it)
__device__ __forceinline__ uint32_t fl_pwin_nproc(const FlPWin& w, uint32_t q, int lane, uint32_t& endmask) {
    const uint32_t i = q + (uint32_t)lane;
    endmask = __ballot_sync(FL_FULL, !w.valid || w.nx != i);
    if (!endmask) return 32u;
    const int el = __ffs((int)endmask) - 1;
    const int el_valid = __shfl_sync(FL_FULL, (int)w.valid, el);
    return (uint32_t)el + (el_valid ? 1u : 0u);
}

// the two serial chains of a window (response time; clamp if max_slope): per-lane terms staged in shared memory,
// every lane runs the identical chain over broadcast reads (8 terms fetched together, then 8 dependent additions).
// Leaves the lane's own results in my_rt / my_z.
__device__ __forceinline__ void fl_pwin_compute(const FlSplit& e, const FlPWin& w, uint32_t q, uint32_t nproc, int lane,
                                                uint32_t root, double& rt_prev, double& z_prev, double e_out,
                                                double rt_out, bool& changed, FlChainSmem& sm, double& my_rt,
                                                double& my_z) {
    __syncwarp();
    sm.in[lane] = ((uint32_t)lane < nproc) ? w.t : 0.0;  // padding: 0.0 + (r + 0.0) == r (r >= +0.0)
    __syncwarp();
    {
        double r = rt_prev;
        for (uint32_t k0 = 0; k0 < nproc; k0 += 8u) {
            double v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = sm.in[k0 + (uint32_t)j];
#pragma unroll
            for (int j = 0; j < 8; ++j) { r = 0.0 + (r + v[j]); v[j] = r; }
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) sm.out[k0 + (uint32_t)j] = v[j];
            }
        }
        rt_prev = r;
    }
    __syncwarp();
    my_rt = sm.out[lane];
    double z = e_out + w.up * fmax(my_rt - rt_out, 0.0);
    if (e.tan_slope) {
        __syncwarp();
        sm.in[lane] = z;
        sm.aux1[lane] = w.ms;
        sm.aux2[lane] = w.d;
        __syncwarp();
        double zp = z_prev;
        for (uint32_t k = 0; k < nproc; ++k) {
            double zk = sm.in[k];
            const double msk = sm.aux1[k];
            const double dk = sm.aux2[k];
            if (msk == msk) {
                const double slope = (zk - zp) / dk;
                if (slope > msk) zk = zp + msk * dk;
            }
            zp = zk;
            sm.out[k] = zk;
        }
        z_prev = zp;
        __syncwarp();
        z = sm.out[lane];
    } else {
        z_prev = __shfl_sync(FL_FULL, z, (int)nproc - 1);
    }
    my_z = z;
    if ((uint32_t)lane < nproc) {
        const uint32_t i = q + (uint32_t)lane;
        changed |= (z != w.eold);
        e.elev[i] = z;
        e.rt[i] = my_rt;
        e.root_of[i] = root;
    }
}

// queue entries for the registered children of a window's sites: one reservation per window.  The reservation (an atomic
// on the queue tail, ~1 us round trip) is ISSUED when the window has been written and CONSUMED one window later
// (fl_push_finish), so the chain of windows of a long run does not wait for it.
#ifndef FL_PUSH_PIPELINE
#define FL_PUSH_PIPELINE 1  // build-time A/B switch (tools/ab_build.py); 0: reserve and fill at once
#endif
struct FlPush {
    uint32_t total, cnt, off, m, s0, site, sbase;
    double rt, z;
};
__device__ __forceinline__ void fl_push_begin(const FlSplit& e, const FlPWin& w, uint32_t q, uint32_t nproc, int lane,
                                              double my_rt, double my_z, FlPush& p) {
    p.m = ((uint32_t)lane < nproc) ? w.pm : 0u;
    p.cnt = (uint32_t)__popc(p.m);
    uint32_t off = p.cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FL_FULL, off, o);
        if (lane >= o) off += v;
    }
    p.total = __shfl_sync(FL_FULL, off, 31);
    p.off = off - p.cnt;
    p.s0 = w.s0; p.site = q + (uint32_t)lane; p.rt = my_rt; p.z = my_z;
    p.sbase = 0u;
    if (p.total && lane == 0) p.sbase = atomicAdd(&e.flags[FLQ_TAIL], p.total);  // (the result is not needed yet)
}
__device__ __forceinline__ void fl_push_finish(const FlSplit& e, FlPush& p, uint32_t root) {
    if (!p.total) return;
    const uint32_t sbase = __shfl_sync(FL_FULL, p.sbase, 0);
    if (p.cnt) {
        uint32_t at = sbase + p.off;
        fl_push_masked(e, p.site, p.s0, p.m, [&](uint32_t c) {
            if (at < e.n) flq_put(e, at, c, root, p.rt, p.z, true);
            else atomicOr(&e.flags[FL_FLAG_BROKEN], 2u);
            ++at;
        });
    }
    p.total = 0u;
}

// a run from q on, by the whole warp; PUSH: registered children become queue entries (k_elev_top)
template <bool PUSH, int DEPTH>
__device__ __forceinline__ void fl_run_warp(const FlSplit& e, uint32_t q, const FlSegStart& s, bool& changed,
                                            FlChainSmem& sm, const FlPWin* first = nullptr) {
    const int lane = threadIdx.x & 31;
    const bool live = s.root != FL_NONE;
    double rt_prev = s.rt_prev, z_prev = s.z_prev;
    FlPWin ring[DEPTH];
    // (`first`: the caller requested the first window before it knew what the run starts from)
    ring[0] = first ? *first : fl_pwin_load<PUSH>(e, q, lane, live);
    uint32_t endmask;
    uint32_t nproc = fl_pwin_nproc(ring[0], q, lane, endmask);
    if (!endmask) {  // the run goes on beyond the first window: keep DEPTH windows in flight
#pragma unroll
        for (int j = 1; j < DEPTH; ++j) ring[j] = fl_pwin_load<PUSH>(e, q + 32u * (uint32_t)j, lane, live);
    }
    FlPush pend;
    pend.total = 0u; pend.cnt = 0u; pend.off = 0u; pend.m = 0u; pend.s0 = 0u; pend.site = 0u; pend.sbase = 0u;
    pend.rt = 0.0; pend.z = 0.0;
    for (;;) {
        const FlPWin cur = ring[0];
        if (!endmask) {
#pragma unroll
            for (int j = 0; j + 1 < DEPTH; ++j) ring[j] = ring[j + 1];
            ring[DEPTH - 1] = fl_pwin_load<PUSH>(e, q + 32u * (uint32_t)DEPTH, lane, live);
        }
        double my_rt = 0.0, my_z = 0.0;
        if (live) fl_pwin_compute(e, cur, q, nproc, lane, s.root, rt_prev, z_prev, s.e_out, s.rt_out, changed, sm, my_rt, my_z);
        else if ((uint32_t)lane < nproc) e.root_of[q + (uint32_t)lane] = FL_NONE;
        if (PUSH) {
            fl_push_finish(e, pend, s.root);  // the previous window's children (its reservation has arrived by now)
            fl_push_begin(e, cur, q, nproc, lane, my_rt, my_z, pend);
            if (!FL_PUSH_PIPELINE) fl_push_finish(e, pend, s.root);
        }
        if (endmask) break;
        q += 32u;
        nproc = fl_pwin_nproc(ring[0], q, lane, endmask);
    }
    if (PUSH) fl_push_finish(e, pend, s.root);
}

// ------------------------------------------------------------------------------------------------
// top: one warp per queue entry
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_elev_top(FlSplit e) {
    __shared__ FlChainSmem chain_smem[8];
    FlChainSmem& sm = chain_smem[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    bool changed = false;
    for (;;) {
        // Tickets are handed out in queue order; the warp waits for the entry with its number.  A waiting warp polls
        // nothing but its own entry's flag word (its own address: no hot spot), and looks at the shared counters only
        // once in a while -- done / tail are the words every producer's atomics go to.
        uint32_t t = 0u;
        if (lane == 0) t = atomicAdd(&e.flags[FLQ_HEAD], 1u);
        t = __shfl_sync(FL_FULL, t, 0);
        if (t >= e.n) break;  // more tickets than entries can ever exist
        uint32_t q = FL_NONE, idle = 0u;
        unsigned long long wait_from = 0ull;
        for (;;) {
            // every lane runs its own acquire load of the flag word (one broadcast transaction)
            const unsigned long long f = flq_ld_acquire(&e.queue[t].flag);
            if ((uint32_t)(f >> 32) == e.epoch) { q = (uint32_t)f; break; }
            ++idle;
            if ((idle & 31u) == 0u) {
                // All work is done when every entry ever appended has been finished: an entry is appended only by an
                // unfinished one, so done == tail is final -- PROVIDED the tail read is not older than the done read and
                // counts every entry the finished ones appended.  Hence: `done` is read with acquire (the tail load
                // cannot be served before it) and incremented with release (a finisher's reservations on `tail` are
                // visible before its increment of `done`).  Two relaxed loads let a warp pair a NEW done with an OLD tail
                // (different L2 slices), leave with a ticket whose entry was about to be written, and orphan that entry:
                // the sweep then waited for it until the limit (profiles/r2v_bench_2gpu_failed.err: tail 278, done 277).
                uint32_t fin = 0u;
                if (lane == 0) {
                    const uint32_t done = fl_ld_acquire(&e.flags[FLQ_DONE]);
                    const uint32_t tail = fl_ld_relaxed(&e.flags[FLQ_TAIL]);
                    fin = (done == tail && t >= tail) ? 1u : 0u;
                }
                if (__shfl_sync(FL_FULL, fin, 0)) break;
                if ((idle & 1023u) == 0u) {  // (the wall clock once per ~1000 polls)
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    now = __shfl_sync(FL_FULL, now, 0);
                    if (wait_from == 0ull) wait_from = now;
                    else if (now - wait_from > FL_TOP_WAIT_LIMIT_NS) { if (lane == 0) atomicOr(&e.flags[FL_FLAG_BROKEN], 2u); break; }
                }
            }
            __nanosleep(idle < 8u ? 32u : 100u);
        }
        if (q == FL_NONE) break;
        // the run's first window is requested together with the entry's payload (one round trip less per hand-off)
        const FlPWin w0 = fl_pwin_load<true>(e, q, lane, true);
        FlSegStart s;
        s.root = e.queue[t].root;
        s.rt_prev = e.queue[t].rt_p;
        s.z_prev = e.queue[t].z_p;
        // the root's values were written by k_elev_plan (an earlier launch)
        s.e_out = s.root != FL_NONE ? e.elev[s.root] : 0.0;
        s.rt_out = s.root != FL_NONE ? e.rt[s.root] : 0.0;
        fl_run_warp<true, FL_PDEPTH>(e, q, s, changed, sm, FL_TOP_EARLY_WINDOW ? &w0 : nullptr);
        __syncwarp();
        if (lane == 0)  // release: this entry's reservations on the queue tail are visible before it counts as done
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&e.flags[FLQ_DONE]) : "memory");
    }
    if (changed) e.flags[FL_FLAG_CHANGED] = 1u;
}
#endif

// ------------------------------------------------------------------------------------------------
// low: all segments of ONE nesting height below the cut, one thread per position
// ------------------------------------------------------------------------------------------------
#ifdef FL_EMU
__global__ void __launch_bounds__(256) k_elev_low(FlSplit e, uint32_t level) {
    // emulation: one thread per position
    const uint32_t h = FL_TID;
    if (h >= e.n || e.hgt[h] != level) return;
    bool changed = false;
    uint32_t q = h;
    FlSegStart s;
    s.root = FL_NONE; s.rt_prev = 0.0; s.z_prev = 0.0; s.e_out = 0.0; s.rt_out = 0.0;
    const uint32_t p = e.recv[h];
    bool ended = false;
    if (p == h) {
        s = fl_root_site(e, h, e.tcel[h], changed);
        q = h + 1u;
        ended = !(q < e.n && e.recv[q] == h);
    } else {  // the receiver's segment is strictly higher: finished by an earlier launch
        s.root = e.root_of[p];
        if (s.root != FL_NONE) {
            s.rt_prev = e.rt[p];
            s.z_prev = e.elev[p];  // the receiver already holds its NEW elevation
            s.e_out = e.elev[s.root];
            s.rt_out = e.rt[s.root];
        }
    }
    while (!ended) ended = fl_run_batch<FL_PB>(e, q, s, changed);
    if (changed) e.flags[FL_FLAG_CHANGED] = 1u;
}
#else
// The heads of this height were listed by k_elev_plan together with their receivers; a persistent grid walks the list
// 256 heads per block at a time.  Phase 1, one thread per head: the first batch's loads and the receiver's values are
// requested together (three dependent round trips per head: list -> receiver + sites -> tree root), runs of up to
// FL_PSHORT sites are finished by the thread.  Longer ones go to a list in shared memory and are walked by whole warps
// in phase 2 -- kept apart so that the window machinery does not sit inside the thread-per-head loop.
#ifndef FL_LOW_MINBLOCKS
#define FL_LOW_MINBLOCKS 4  // resident blocks per SM the register allocation aims at
#endif
#ifndef FL_PB_LOW
#define FL_PB_LOW 2  // sites per batch of a thread of k_elev_low (measured: 2 with 4 resident blocks beats 4 with 3, r2m_ab.txt)
#endif
struct FlLongRun { uint32_t q, root; double rt_prev, z_prev, e_out, rt_out; };
__global__ void __launch_bounds__(256, FL_LOW_MINBLOCKS) k_elev_low(FlSplit e, uint32_t level, size_t region) {
    __shared__ FlChainSmem chain_smem[8];
    __shared__ FlLongRun longs[256];
    __shared__ uint32_t n_long;
    const uint32_t total = e.flags[FLQ_LOW + level];
    const uint2* __restrict__ list = e.low_list + region;  // = fl_low_region(e.n, level)
    bool changed = false;
    for (unsigned long long g0 = (unsigned long long)blockIdx.x * 256u; g0 < total; g0 += (unsigned long long)gridDim.x * 256u) {
        if (threadIdx.x == 0u) n_long = 0u;
        __syncthreads();
        const unsigned long long k = g0 + threadIdx.x;
        if (k < total) {
            const uint2 hp = list[k];
            const uint32_t h = hp.x, p = hp.y;
            uint32_t q = h;
            FlSegStart s;
            s.root = FL_NONE; s.rt_prev = 0.0; s.z_prev = 0.0; s.e_out = 0.0; s.rt_out = 0.0;
            bool ended = false;
            if (p == h) {  // a tree root below the cut (small trees)
                s = fl_root_site(e, h, e.tcel[h], changed);
                q = h + 1u;
                ended = !(q < e.n && e.recv[q] == h);
                if (!ended) ended = fl_run_batch<FL_PB_LOW>(e, q, s, changed);
            } else {  // the receiver's segment is strictly higher: finished by an earlier launch
                FlBatch<FL_PB_LOW> b;
                fl_batch_load<FL_PB_LOW>(e, q, true, b);
                const uint32_t root_p = e.root_of[p];
                const double rt_p = e.rt[p];    // (rt of a tree without outlet is never used)
                const double z_p = e.elev[p];   // the receiver already holds its NEW elevation
                s.root = root_p;
                if (root_p != FL_NONE) {
                    s.rt_prev = rt_p;
                    s.z_prev = z_p;
                    s.e_out = e.elev[root_p];
                    s.rt_out = e.rt[root_p];
                }
                ended = fl_batch_compute<FL_PB_LOW>(e, q, s, changed, b);
            }
#pragma unroll 1
            for (int r = 1; r < FL_PSHORT / FL_PB_LOW && !ended; ++r) ended = fl_run_batch<FL_PB_LOW>(e, q, s, changed);
            if (!ended) {
                FlLongRun& r = longs[atomicAdd(&n_long, 1u)];  // (at most one per thread of the block)
                r.q = q; r.root = s.root; r.rt_prev = s.rt_prev; r.z_prev = s.z_prev; r.e_out = s.e_out; r.rt_out = s.rt_out;
            }
        }
        __syncthreads();
        const uint32_t nl = n_long;
        for (uint32_t j = threadIdx.x >> 5; j < nl; j += 8u) {  // phase 2: a warp per long run
            FlSegStart w;
            w.root = longs[j].root; w.rt_prev = longs[j].rt_prev; w.z_prev = longs[j].z_prev;
            w.e_out = longs[j].e_out; w.rt_out = longs[j].rt_out;
            fl_run_warp<false, 1>(e, longs[j].q, w, changed, chain_smem[threadIdx.x >> 5]);  // (rare: one window ahead is enough)
        }
        __syncthreads();
    }
    if (changed) e.flags[FL_FLAG_CHANGED] = 1u;
}
#endif
