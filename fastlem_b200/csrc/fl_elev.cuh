// fl_elev.cuh -- K5, response time + elevation + max_slope clamp + `changed` (generator.rs:162-203), as ONE launch.
//
// The sweep runs outlet -> upstream.  A site needs nothing but its receiver's new values, so the work is handed
// DOWN the forest through a queue of segment heads (a segment = a run of positions q, q+1, ... with recv[q+1] == q,
// fl_flow.cuh): whoever has written site p pushes every child of p that starts a segment of its own.  No ordering of
// the segments by nesting height, no launch per level, no level offsets on the host -- the queue order IS a
// topological order.  Every floating-point operation is the reference's, in the reference's order:
//     celerity = k_i * A_i^0.5 ;  rt_i = 0.0 + (rt_recv + 1.0 / celerity * d_i)                 generator.rs:162-174
//     z = e_outlet + u_i * max(rt_i - rt_outlet, 0.0) ; clamp against the receiver's NEW elevation   :177-203
//
// Queues (entries are (epoch << 32 | site); epoch = number of this launch, so the arrays are never cleared):
//   short queue : segment heads.  Lanes of the "short" warps hold one ticket (slot number) each, poll their slot,
//                 walk up to FL_PSHORT sites of the segment alone and push the children of those sites;
//   long queue  : the rest of a segment that went on beyond FL_PSHORT sites.  "Long" warps take one ticket per warp
//                 and walk the rest in 32-site windows (coalesced loads, FL_PDEPTH windows prefetched, the serial
//                 additions staged through shared memory), pushing the children window by window.
// Seeds are the outlets (an outlet is its own receiver, hence always a segment head; trees without an outlet are
// never visited: generator.rs:149).  `pending` = entries pushed and not yet finished; a warp that finds nothing
// ready leaves when it reads pending == 0.
// Ordering between threads: producer = result stores, fence.acq_rel.gpu, st.relaxed entry; consumer = ld.relaxed
// entry, fence.acq_rel.gpu, loads (fl_flow.cuh, release / acquire patterns of the PTX memory model).
#pragma once
#include "fl_flow.cuh"

#define FL_PB 4       // sites per thread-level batch (loads issued together)
#define FL_PSHORT 8   // sites a lane walks alone before the rest of the segment goes to the long queue
#define FL_PDEPTH 3   // windows of a long segment kept in flight

enum { FLQ_TAIL = 16, FLQ_HEAD = 17, FLQ_LTAIL = 18, FLQ_LHEAD = 19, FLQ_PENDING = 20 };  // words of d_flags

struct FlPush {
    uint32_t n;
    const uint32_t* row_ptr;
    const uint32_t* col;
    const uint32_t* recv;
    const uint32_t* cmask;
    const double* drecv;
    const double* erod;
    const double* A;
    const double* uplift;
    const double* tan_slope;  // may be null
    double* elev;
    double* rt;
    uint32_t* root_of;
    unsigned long long* queue;   // n entries
    unsigned long long* lqueue;  // lcap entries
    uint32_t lcap;
    uint32_t epoch;
    const uint32_t* seeds;  // the outlets in the current numbering
    uint32_t n_seeds;
    uint32_t* flags;
};

__device__ __forceinline__ unsigned long long flq_entry(uint32_t epoch, uint32_t site) {
    return ((unsigned long long)epoch << 32) | (unsigned long long)site;
}
#ifdef FL_EMU
__device__ __forceinline__ unsigned long long flq_ld(const unsigned long long* p) { return *p; }
__device__ __forceinline__ void flq_st(unsigned long long* p, unsigned long long v) { *p = v; }
#else
__device__ __forceinline__ unsigned long long flq_ld(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void flq_st(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#endif

// seeds + counters (before the sweep; one small launch)
__global__ void __launch_bounds__(256) k_push_seed(FlPush e) {
    const uint32_t k = FL_TID;
    if (k == 0u) {
        e.flags[FLQ_TAIL] = e.n_seeds;
        e.flags[FLQ_PENDING] = e.n_seeds;
        e.flags[FLQ_HEAD] = 0u; e.flags[FLQ_LTAIL] = 0u; e.flags[FLQ_LHEAD] = 0u;
    }
    if (k < e.n_seeds) e.queue[k] = flq_entry(e.epoch, e.seeds[k]);
}

// what a segment starts from: the values of the head's receiver (another segment, already written) or, for a
// tree root (an outlet), the root's own old elevation
struct FlSegStart {
    uint32_t root;
    double rt_prev, z_prev, e_out, rt_out;
    bool is_root;
};
__device__ __forceinline__ FlSegStart fl_push_start(const FlPush& e, uint32_t h) {
    FlSegStart s;
    const uint32_t p = e.recv[h];
    s.is_root = (p == h);
    if (s.is_root) {
        s.root = h;
        s.rt_prev = 0.0;
        s.z_prev = e.elev[h];  // has_edge(i,i) is false: the clamp compares with the site's own old elevation
        s.e_out = s.z_prev;
        s.rt_out = 0.0;
    } else {  // written by another thread of this launch: read through L2
        s.root = fl_ld_cg(&e.root_of[p]);
        s.rt_prev = fl_ld_cg(&e.rt[p]);
        s.z_prev = fl_ld_cg(&e.elev[p]);  // the receiver already holds its NEW elevation
        s.e_out = fl_ld_cg(&e.elev[s.root]);
        s.rt_out = fl_ld_cg(&e.rt[s.root]);
    }
    return s;
}

// up to B sites of one segment by one thread, starting at q; returns true when the segment ended inside the batch.
// nchild counts the children that start segments of their own (every child except the chain child q+1).
template <int B>
__device__ __forceinline__ bool fl_push_batch(const FlPush& e, uint32_t& q, uint32_t h, FlSegStart& s, bool& changed,
                                              uint32_t& nchild) {
    const uint32_t nb = e.n - q < (uint32_t)B ? e.n - q : (uint32_t)B;
    double d[B], t[B], up[B], eo[B], ms[B];
    uint32_t nx[B], cm[B];
#pragma unroll
    for (int k = 0; k < B; ++k) {
        d[k] = 1.0; t[k] = 0.0; up[k] = 0.0; eo[k] = 0.0; ms[k] = 0.0; nx[k] = FL_NONE; cm[k] = 0u;
        if ((uint32_t)k < nb) {
            const uint32_t i = q + k;
            d[k] = e.drecv[i];
            const double celerity = e.erod[i] * sqrt(e.A[i]);
            t[k] = 1.0 / celerity * d[k];
            up[k] = e.uplift[i];
            eo[k] = e.elev[i];
            if (e.tan_slope) ms[k] = e.tan_slope[i];
            nx[k] = (i + 1u < e.n) ? e.recv[i + 1u] : FL_NONE;
            cm[k] = e.cmask[i];
        }
    }
    bool ended = false;
    uint32_t walked = 0u;
#pragma unroll
    for (int k = 0; k < B; ++k) {
        if (!ended && (uint32_t)k < nb) {
            const uint32_t i = q + k;
            ++walked;
            const double rti = 0.0 + (s.rt_prev + t[k]);
            if (s.is_root && i == h) s.rt_out = rti;
            double z = s.e_out + up[k] * fmax(rti - s.rt_out, 0.0);
            if (e.tan_slope) {
                if (ms[k] == ms[k]) {  // not NaN: Some(max_slope)
                    const double slope = (z - s.z_prev) / d[k];
                    if (slope > ms[k]) z = s.z_prev + ms[k] * d[k];
                }
            }
            changed |= (z != eo[k]);
            if (s.is_root && i == h) s.e_out = z;  // later sites read elevations[outlet] after the outlet's own update
            e.elev[i] = z;
            e.rt[i] = rti;
            e.root_of[i] = s.root;
            s.rt_prev = rti;
            s.z_prev = z;
            const bool chain = nx[k] == i;
            nchild += (uint32_t)__popc(cm[k]) - (chain ? 1u : 0u);
            if (!chain) ended = true;
        }
    }
    q += walked;  // one past the last site written
    return ended || q >= e.n;
}

// entries for the children of the sites [first, end) that start segments of their own, written from slot `at` on
template <class F>
__device__ __forceinline__ void fl_push_children(const FlPush& e, uint32_t first, uint32_t end, F&& put) {
    for (uint32_t i = first; i < end; ++i) {
        uint32_t m = e.cmask[i];
        if (!m) continue;
        const uint32_t s0 = e.row_ptr[i];
        const bool chain = (i + 1u < e.n) && (e.recv[i + 1u] == i);
        while (m) {
            const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
            m &= m - 1u;
            const uint32_t c = e.col[s0 + b];
            if (chain && c == i + 1u) continue;
            put(c);
        }
    }
}

#ifdef FL_EMU
// host emulation: one thread drains the queue in FIFO order (parents before children)
__global__ void k_elev_push(FlPush e) {
    if (FL_TID != 0u) return;
    uint32_t head = 0, tail = e.flags[FLQ_TAIL];
    bool changed = false;
    while (head < tail) {
        const unsigned long long ent = e.queue[head++];
        if ((uint32_t)(ent >> 32) != e.epoch) { atomicOr(&e.flags[FL_FLAG_BROKEN], 2u); return; }
        const uint32_t h = (uint32_t)ent;
        FlSegStart s = fl_push_start(e, h);
        uint32_t q = h, nchild = 0;
        while (!fl_push_batch<FL_PB>(e, q, h, s, changed, nchild)) {}
        uint32_t pushed = 0;  // the segment's sites are [h, q)
        fl_push_children(e, h, q, [&](uint32_t c) { e.queue[tail++] = flq_entry(e.epoch, c); ++pushed; });
        if (pushed != nchild) { atomicOr(&e.flags[FL_FLAG_BROKEN], 2u); return; }
    }
    e.flags[FLQ_TAIL] = tail;
    e.flags[FLQ_HEAD] = head;
    e.flags[FLQ_PENDING] = 0u;
    if (changed) e.flags[FL_FLAG_CHANGED] = 1u;
}
#else

// ------------------------------------------------------------------------------------------------
// short warps: one ticket per lane
// ------------------------------------------------------------------------------------------------
__device__ void fl_push_short_warp(const FlPush& e) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t my = FL_NONE;
    bool changed = false;
    uint32_t idle = 0u;
    for (;;) {
        const uint32_t need = __ballot_sync(FL_FULL, my == FL_NONE);
        if (need) {
            uint32_t base = 0u;
            if (lane == 0) base = atomicAdd(&e.flags[FLQ_HEAD], (uint32_t)__popc(need));
            base = __shfl_sync(FL_FULL, base, 0);
            if (my == FL_NONE) my = base + (uint32_t)__popc(need & lt_mask);
        }
        unsigned long long ent = 0ull;
        if (my < e.n) ent = flq_ld(&e.queue[my]);
        const bool ready = (uint32_t)(ent >> 32) == e.epoch;
        const uint32_t rmask = __ballot_sync(FL_FULL, ready);
        if (!rmask) {
            uint32_t pend = 1u;
            if (lane == 0) pend = fl_ld_relaxed(&e.flags[FLQ_PENDING]);
            pend = __shfl_sync(FL_FULL, pend, 0);
            if (pend == 0u) break;
            const uint32_t ns = 64u << (idle < 5u ? idle : 5u);
            __nanosleep(ns);
            ++idle;
            if (idle > (1u << 22)) { if (lane == 0) atomicOr(&e.flags[FL_FLAG_BROKEN], 2u); break; }
            continue;
        }
        idle = 0u;
        fl_fence_acquire();  // acquire side of the producers' fence + entry store
        uint32_t nchild = 0u, cont = FL_NONE, h = 0u, q = 0u;
        if (ready) {
            h = (uint32_t)ent;
            q = h;
            FlSegStart s = fl_push_start(e, h);
            bool ended = fl_push_batch<FL_PB>(e, q, h, s, changed, nchild);
#pragma unroll
            for (int r = 1; r < FL_PSHORT / FL_PB; ++r)
                if (!ended) ended = fl_push_batch<FL_PB>(e, q, h, s, changed, nchild);
            if (!ended) cont = q;  // the rest of a long segment: a warp continues from q (its receiver is q - 1)
        }
        // one reservation per warp and round
        uint32_t off = nchild;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t w = __shfl_up_sync(FL_FULL, off, o);
            if (lane >= o) off += w;
        }
        const uint32_t total = __shfl_sync(FL_FULL, off, 31);
        off -= nchild;
        const uint32_t lmask = __ballot_sync(FL_FULL, cont != FL_NONE);
        const uint32_t ltotal = (uint32_t)__popc(lmask);
        if (total | ltotal) {
            uint32_t sbase = 0u, lbase = 0u;
            if (lane == 0) {
                atomicAdd(&e.flags[FLQ_PENDING], total + ltotal);  // before the entries can be seen
                if (total) sbase = atomicAdd(&e.flags[FLQ_TAIL], total);
                if (ltotal) lbase = atomicAdd(&e.flags[FLQ_LTAIL], ltotal);
            }
            sbase = __shfl_sync(FL_FULL, sbase, 0);
            lbase = __shfl_sync(FL_FULL, lbase, 0);
            __syncwarp();         // lane 0's increment of `pending` happens before every lane's fence
            fl_fence_release();   // each lane: its results (and, by cumulativity, the increment) before its entries
            if (ready) {
                uint32_t at = sbase + off;
                fl_push_children(e, h, q, [&](uint32_t c) { flq_st(&e.queue[at++], flq_entry(e.epoch, c)); });
                if (cont != FL_NONE) {
                    const uint32_t slot = lbase + (uint32_t)__popc(lmask & lt_mask);
                    if (slot < e.lcap) flq_st(&e.lqueue[slot], flq_entry(e.epoch, cont));
                    else atomicOr(&e.flags[FL_FLAG_BROKEN], 2u);
                }
            }
        }
        if (lane == 0) atomicAdd(&e.flags[FLQ_PENDING], 0u - (uint32_t)__popc(rmask));
        if (ready) my = FL_NONE;
    }
    if (changed) e.flags[FL_FLAG_CHANGED] = 1u;
}

// ------------------------------------------------------------------------------------------------
// long warps: one ticket per warp; the segment from q on in 32-site windows
// ------------------------------------------------------------------------------------------------
struct FlPWin {  // one window: lane l holds site base + l
    double t, up, eold, ms, d;
    uint32_t nx, cm, s0;
    bool valid;
};

__device__ __forceinline__ FlPWin fl_pwin_load(const FlPush& e, uint32_t base, int lane) {
    FlPWin w;
    w.t = 0.0; w.up = 0.0; w.eold = 0.0; w.ms = 0.0; w.d = 1.0; w.nx = FL_NONE; w.cm = 0u; w.s0 = 0u;
    const unsigned long long i64 = (unsigned long long)base + (unsigned)lane;
    w.valid = i64 < e.n;
    if (w.valid) {
        const uint32_t i = (uint32_t)i64;
        w.d = e.drecv[i];
        const double celerity = e.erod[i] * sqrt(e.A[i]);
        w.t = 1.0 / celerity * w.d;
        w.up = e.uplift[i];
        w.eold = e.elev[i];
        w.nx = (i + 1u < e.n) ? e.recv[i + 1u] : FL_NONE;
        w.cm = e.cmask[i];
        w.s0 = e.row_ptr[i];
        if (e.tan_slope) w.ms = e.tan_slope[i];
    }
    return w;
}

__device__ __forceinline__ uint32_t fl_pwin_nproc(const FlPWin& w, uint32_t q, int lane, uint32_t& endmask) {
    const uint32_t i = q + (uint32_t)lane;
    endmask = __ballot_sync(FL_FULL, !w.valid || w.nx != i);
    if (!endmask) return 32u;
    const int el = __ffs((int)endmask) - 1;
    const int el_valid = __shfl_sync(FL_FULL, (int)w.valid, el);
    return (uint32_t)el + (el_valid ? 1u : 0u);
}

// the two serial chains of a window (response time; clamp if max_slope): per-lane terms staged in shared memory,
// every lane runs the identical chain over broadcast reads (8 terms fetched together, then 8 dependent additions)
__device__ __forceinline__ void fl_pwin_compute(const FlPush& e, const FlPWin& w, uint32_t q, uint32_t nproc, int lane,
                                                uint32_t root, double& rt_prev, double& z_prev, double e_out,
                                                double rt_out, bool& changed, FlChainSmem& sm) {
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
    const double my_rt = sm.out[lane];
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
    if ((uint32_t)lane < nproc) {
        const uint32_t i = q + (uint32_t)lane;
        changed |= (z != w.eold);
        e.elev[i] = z;
        e.rt[i] = my_rt;
        e.root_of[i] = root;
    }
}

// children of a window's sites: one reservation per window
__device__ __forceinline__ void fl_pwin_push(const FlPush& e, const FlPWin& w, uint32_t q, uint32_t nproc, int lane) {
    const uint32_t i = q + (uint32_t)lane;
    const bool inwin = (uint32_t)lane < nproc;
    const bool chain = inwin && w.nx == i;
    const uint32_t cnt = inwin ? (uint32_t)__popc(w.cm) - (chain ? 1u : 0u) : 0u;
    uint32_t off = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FL_FULL, off, o);
        if (lane >= o) off += v;
    }
    const uint32_t total = __shfl_sync(FL_FULL, off, 31);
    if (!total) return;
    off -= cnt;
    // the children's ids first (independent loads), while the reservation is under way
    uint32_t kid[4];
    uint32_t m = w.cm;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        kid[k] = FL_NONE;
        if (cnt && m) {
            const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
            m &= m - 1u;
            kid[k] = e.col[w.s0 + b];
        }
    }
    uint32_t sbase = 0u;
    if (lane == 0) {
        atomicAdd(&e.flags[FLQ_PENDING], total);
        sbase = atomicAdd(&e.flags[FLQ_TAIL], total);
    }
    sbase = __shfl_sync(FL_FULL, sbase, 0);
    __syncwarp();
    fl_fence_release();
    if (cnt) {
        uint32_t at = sbase + off;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (kid[k] != FL_NONE && !(chain && kid[k] == i + 1u)) flq_st(&e.queue[at++], flq_entry(e.epoch, kid[k]));
        while (m) {  // more than four children
            const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
            m &= m - 1u;
            const uint32_t c = e.col[w.s0 + b];
            if (chain && c == i + 1u) continue;
            flq_st(&e.queue[at++], flq_entry(e.epoch, c));
        }
    }
}

__device__ void fl_push_long_warp(const FlPush& e, FlChainSmem& sm) {
    const int lane = threadIdx.x & 31;
    bool changed = false;
    for (;;) {
        uint32_t t = 0u;
        if (lane == 0) t = atomicAdd(&e.flags[FLQ_LHEAD], 1u);
        t = __shfl_sync(FL_FULL, t, 0);
        uint32_t q = FL_NONE, idle = 0u;
        for (;;) {
            uint32_t got = FL_NONE, pend = 1u;
            if (lane == 0) {
                if (t < e.lcap) {
                    const unsigned long long ent = flq_ld(&e.lqueue[t]);
                    if ((uint32_t)(ent >> 32) == e.epoch) got = (uint32_t)ent;
                }
                if (got == FL_NONE) pend = fl_ld_relaxed(&e.flags[FLQ_PENDING]);
            }
            got = __shfl_sync(FL_FULL, got, 0);
            pend = __shfl_sync(FL_FULL, pend, 0);
            if (got != FL_NONE) { q = got; break; }
            if (pend == 0u) break;
            const uint32_t ns = 64u << (idle < 5u ? idle : 5u);
            __nanosleep(ns);
            ++idle;
            if (idle > (1u << 22)) { if (lane == 0) atomicOr(&e.flags[FL_FLAG_BROKEN], 2u); break; }
        }
        if (q == FL_NONE) break;
        fl_fence_acquire();  // every lane: acquire side of the producer's fence + entry store (lane 0 read the entry;
        __syncwarp();        // its fence and this barrier order the other lanes' loads behind it)
        // q continues a segment: its receiver q - 1 holds the running values
        FlSegStart s = fl_push_start(e, q);
        double rt_prev = s.rt_prev, z_prev = s.z_prev;
        FlPWin ring[FL_PDEPTH];
#pragma unroll
        for (int j = 0; j < FL_PDEPTH; ++j) ring[j] = fl_pwin_load(e, q + 32u * (uint32_t)j, lane);
        for (;;) {
            uint32_t endmask;
            const uint32_t nproc = fl_pwin_nproc(ring[0], q, lane, endmask);
            const FlPWin cur = ring[0];
#pragma unroll
            for (int j = 0; j + 1 < FL_PDEPTH; ++j) ring[j] = ring[j + 1];
            if (!endmask) ring[FL_PDEPTH - 1] = fl_pwin_load(e, q + 32u * (uint32_t)FL_PDEPTH, lane);
            fl_pwin_compute(e, cur, q, nproc, lane, s.root, rt_prev, z_prev, s.e_out, s.rt_out, changed, sm);
            fl_pwin_push(e, cur, q, nproc, lane);
            if (endmask) break;
            q += 32u;
        }
        if (lane == 0) atomicAdd(&e.flags[FLQ_PENDING], 0u - 1u);
    }
    if (changed) e.flags[FL_FLAG_CHANGED] = 1u;
}

// warps 0..3 of a block serve the short queue, warps 4..7 the long queue
__global__ void __launch_bounds__(256) k_elev_push(FlPush e) {
    __shared__ FlChainSmem chain_smem[4];
    const uint32_t w = threadIdx.x >> 5;
    if (w < 4u) fl_push_short_warp(e);
    else fl_push_long_warp(e, chain_smem[w - 4u]);
}
#endif
