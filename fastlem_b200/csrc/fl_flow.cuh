// fl_flow.cuh -- "sweep" = 3: dataflow tree sweeps on DYNAMIC segments (the default).
//
// A segment is a maximal run of positions q, q+1, ... with recv[q+1] == q: a piece of a flow path that is
// contiguous in the current site numbering.  Right after a layout rebuild (fl_paths.cuh, heavy paths) the
// segments are the heavy paths; when receivers change in later iterations a chain simply breaks into
// shorter segments -- nothing has to be patched, so a numbering is kept for many iterations and rebuilt
// only when it has degraded.  Every floating-point addition below is one the reference performs, in the
// reference's order (generator.rs:154-174); only WHO performs it and WHEN is reorganised.
//
// K4 drainage area, bottom-up, no level structure and no waiting:
//   * Leaves need no work: their area is their own cell area, their parent reads it directly.
//   * A site whose non-chain children are all leaves is "simple": its partial sums (pre = a_p + children
//     before the chain child, post1/post2 = children after it, reverse adjacency order) are gathered in a
//     parallel pre-pass (k_simple_pre).
//   * Every other site waits for reports: a flow that finishes a segment head h reports to p = recv[h]; the LAST
//     child to report gathers p's partial sums and publishes them.
//   * A SEGMENT is climbed exactly once, from its tail, when every waiting site on it has been published: the
//     thread that publishes the last one starts the climb (per-segment counter).  So a climb never meets an
//     unfinished site, nobody blocks, nobody spins, and the dependent chain is
//     (nesting depth of the segment forest) x (one climb + one hand-off).
//   * pass 1 (k_area_flow) runs the flows with one thread each.  A climb that has covered `park_after` sites is on
//     a long chain: the thread parks it.  pass 2 (k_area_flow_long) continues the parked climbs with whole
//     warps: 32-site windows fetched by the lanes together, prefetched FL_WDEPTH windows ahead, the serial
//     additions staged through shared memory.  A warp stays a warp through the later hand-offs of its flow.
//   The pass also yields, per segment head, the nesting height (longest chain of hand-offs below it), which
//   orders the top-down sweep exactly for the CURRENT forest.
// K5 response time / elevation, top-down: one launch per nesting height; levels with few segments run one warp
//   per segment, populous levels one thread per segment with the warp finishing the long ones.
#pragma once
#include "fl_paths.cuh"

// state word of a site: [0,8) children reported, [8,12) np (posts; 15 = more than two), [12,30) hpre (max child
// height + 1; 0x3FFFF = too large, read hpre[]), bit 30 partial sums published
#define FL_ST_COUNT_MASK 0x000000FFu
#define FL_ST_NP_SHIFT 8
#define FL_ST_NP_MASK 0x00000F00u
#define FL_ST_HP_SHIFT 12
#define FL_ST_HP_MASK 0x3FFFF000u
#define FL_ST_HP_OVER 0x3FFFFu
#define FL_ST_PRE_READY 0x40000000u
#define FL_TB 4  // sites per batch of a thread-level climb

// optional event counters (build with -DFL_FLOW_STATS; read back through fastlem_debug_fetch stage 9)
#ifdef FL_FLOW_STATS
#define FL_COUNT(f, k, v) atomicAdd(&(f).stats[k], (unsigned long long)(v))
#else
#define FL_COUNT(f, k, v) ((void)0)
#endif
enum { FLS_T_CLIMBS = 0, FLS_T_BATCH, FLS_T_SITES, FLS_T_HEADS, FLS_T_LAST, FLS_T_SEGSTART, FLS_T_PARKED, FLS_T_NOTREADY,
       FLS_W_FLOWS, FLS_W_WINDOWS, FLS_W_SITES, FLS_W_HEADS, FLS_W_LAST, FLS_W_SEGSTART, FLS_W_REDO, FLS_W_NOTREADY,
       FLS_W_CYC_FLOW, FLS_W_CYC_REPORT, FLS_W_CYC_FIRSTWIN, FLS_W_CYC_WIN, FLS_COUNT_N = 24 };
#if defined(FL_FLOW_STATS) && !defined(FL_EMU)
#define FL_CLOCK() clock64()
__device__ __forceinline__ unsigned long long fl_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// timeline of a segment (debug): slot 0 climb started, 1 parked, 2 resumed by a warp, 3 head finished
#define FL_TLOG(f, site, slot) do { if ((f).tlog) (f).tlog[4ull * (f).seg_head[(site)] + (slot)] = fl_gtime(); } while (0)
#else
#define FL_CLOCK() 0ll
#define FL_TLOG(f, site, slot) ((void)0)
#endif

// Hand-offs between threads of one launch follow the PTX memory model's release / acquire patterns at gpu scope:
//   producer:  plain stores ... fence.acq_rel.gpu ; atomic (relaxed)        -- release pattern
//   consumer:  atomic (relaxed) or ld.acquire.gpu ... fence.acq_rel.gpu ; plain loads   -- acquire pattern
// The consumer-side fence sits on the WINNER path only (the last reporter / last publisher / the waiter whose flag
// came up), so the common "somebody else arrives last" path pays nothing.  -DFL_RELAXED_READERS=1 builds the
// round-1 variant without the consumer-side fences (ordering by an address dependency on the atomic's result only;
// not covered by the memory model) for A/B timing: tools/ab_acquire.py, profiles/r2_ab_acquire.txt.
#ifndef FL_RELAXED_READERS
#define FL_RELAXED_READERS 0
#endif
#ifdef FL_EMU
template <class T> __device__ __forceinline__ T fl_ld_cg(const T* p) { return *p; }
__device__ __forceinline__ uint32_t fl_ld_relaxed(const uint32_t* p) { return *p; }
__device__ __forceinline__ uint32_t fl_ld_acquire(const uint32_t* p) { return *p; }
__device__ __forceinline__ uint32_t fl_dep0(uint32_t) { return 0u; }
__device__ __forceinline__ void fl_fence_acquire() {}
#else
template <class T> __device__ __forceinline__ T fl_ld_cg(const T* p) { return __ldcg(p); }
// relaxed gpu-scope load of a flag word (served by L2, pipelinable)
__device__ __forceinline__ uint32_t fl_ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#if FL_RELAXED_READERS
__device__ __forceinline__ uint32_t fl_ld_acquire(const uint32_t* p) { return fl_ld_relaxed(p); }
__device__ __forceinline__ void fl_fence_acquire() {}
// An opaque zero derived from v: makes the following loads ADDRESS-DEPENDENT on the atomic result v (A/B variant only)
__device__ __forceinline__ uint32_t fl_dep0(uint32_t v) {
    uint32_t z;
    asm volatile("and.b32 %0, %1, 0;" : "=r"(z) : "r"(v));
    return z;
}
#else
// acquire load of a flag word at gpu scope: loads that follow it in program order see everything the writer released
__device__ __forceinline__ uint32_t fl_ld_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// acquire side of a hand-off decided by a relaxed atomic: fence after the atomic, before the dependent loads
__device__ __forceinline__ void fl_fence_acquire() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ uint32_t fl_dep0(uint32_t) { return 0u; }
#endif
#endif

__device__ __forceinline__ uint32_t fl_st_publish(uint32_t np, uint32_t hp) {
    const uint32_t h18 = hp < FL_ST_HP_OVER ? hp : FL_ST_HP_OVER;
    return FL_ST_PRE_READY | (np << FL_ST_NP_SHIFT) | (h18 << FL_ST_HP_SHIFT);
}
__device__ __forceinline__ uint32_t fl_st_np(uint32_t s) { return (s & FL_ST_NP_MASK) >> FL_ST_NP_SHIFT; }
__device__ __forceinline__ uint32_t fl_st_hp(uint32_t s, const uint32_t* hpre, uint32_t i) {
    const uint32_t h = (s & FL_ST_HP_MASK) >> FL_ST_HP_SHIFT;
    return h != FL_ST_HP_OVER ? h : fl_ld_cg(&hpre[i]);
}

struct FlFlow {
    uint32_t n;
    const uint32_t* row_ptr;
    const uint32_t* col;
    const uint32_t* recv;
    const uint32_t* cmask;
    const double* areas;
    double* A;
    uint32_t* nwait;           // children that REPORT to a site = its non-leaf segment heads (k_count_waits);
                               // incremental mode: its DIRTY segment heads, bit 31 = site is on the regather list
    const uint32_t* seg_head;  // head (lowest position) of the segment each site is on (max-scan)
    const uint32_t* seg_tail;  // at a head: the tail (highest position) of its segment (full mode)
    uint32_t* seg_wait;        // at a head: number of waiting sites (nwait > 0) on the segment
    uint32_t* seg_done;        // at a head: how many of them have been published (zeroed before the launch)
    uint32_t* state;           // per site, zeroed before the launch (layout above)
    double* pre;
    double* post1;
    double* post2;
    double* xpost;  // FL_XPOST per site: the 3rd .. 8th child after the chain child (np in 3..8)
    uint32_t* hpre;
    double* xbuf;        // running area of a parked climb
    uint32_t* hbuf;      // running nesting height of a parked climb
    uint32_t* hgt;       // out: nesting height for segment heads, FL_NONE elsewhere
    uint32_t* hsuf;      // out: running nesting height of the climb after each site (max over the chain above it)
    // incremental mode (dirty_from != nullptr): only the sites whose drainage area can have changed are redone
    uint32_t* dirty_from;  // at a head: 1 + highest dirty position of the segment (0 = clean)
    uint32_t* rlist;       // sites that regather their partial sums; count in flags[FL_FLAG_NREGATHER]
    uint32_t* slist;       // heads of the dirty segments; count in flags[FL_FLAG_NDIRTY]
    uint32_t* flags;
    uint32_t* parked;    // sites where a long climb was parked for the warp-level pass
    uint32_t* counters;  // [0] = number of parked climbs, [1] = next one to take
    uint32_t park_after; // a thread parks its climb after this many sites (0 = never)
    unsigned long long* stats;
    unsigned long long* tlog;  // FL_FLOW_STATS builds: 4 time stamps per segment head
};

#define FL_NW_RFLAG 0x80000000u
#define FL_NW_COUNT 0x7FFFFFFFu

// Where the climb of a segment starts and what it starts with.  Full mode: at the tail, with nothing.
// Incremental mode: at the highest dirty site; the chain above it is clean, so its area and running
// nesting height are the ones stored by an earlier iteration.
__device__ __forceinline__ uint32_t fl_seg_start(const FlFlow& f, uint32_t sh) {
    return f.dirty_from ? (f.dirty_from[sh] - 1u) : f.seg_tail[sh];
}
__device__ __forceinline__ void fl_seg_entry(const FlFlow& f, uint32_t cur, double& x, uint32_t& hrun, bool& has_chain) {
    x = 0.0; hrun = 0u; has_chain = false;
    if (f.dirty_from && cur + 1u < f.n && f.recv[cur + 1u] == cur) {
        has_chain = true;
        x = f.A[cur + 1u];
        hrun = f.hsuf[cur + 1u];
    }
}

// per-warp staging area of the serial chains (warp-level scans)
#define FL_MAXPOST 8  // children after the chain child kept with a site's partial sums (more: slow path, np = 15)
#define FL_XPOST (FL_MAXPOST - 2)
struct FlChainSmem { double in[32]; double out[32]; double aux1[32]; double aux2[32]; };  // K5 windows
// K4 windows: the additions of a 32-site window, flattened into ONE sequence of terms (per site: its partial sum,
// then each child after the chain child); the prefix sums overwrite the terms in place.
struct FlAreaSmem {
    double t[32 * (1 + FL_MAXPOST) + 8];  // terms
    double p[32 * (1 + FL_MAXPOST) + 8];  // prefix sums
};

// Partial sums of site p over its non-chain children, reverse adjacency order; returns np (15 = more than two
// children after the chain child).  A light child is always a segment head; leaf heads hold A = their cell area
// and height 0 from the pre-pass (k_count_waits) or from their own climb.  `dep` is an opaque zero that orders
// the loads of the children's results after the atomic that made the caller the last reporter.
__device__ __forceinline__ uint32_t fl_gather_lights(const FlFlow& f, uint32_t p, bool has_chain, uint32_t dep,
                                                     double& pre, double& p1, double& p2, uint32_t& hmax) {
    double* const xp = f.xpost + (size_t)p * FL_XPOST;
    pre = f.areas[p];
    p1 = 0.0; p2 = 0.0;
    uint32_t np = 0;
    bool seen = false;
    hmax = 0;
    const uint32_t s0 = f.row_ptr[p];
    uint32_t m = f.cmask[p];
    // all children's ids first (independent loads), then their values, then the ordered additions
    uint32_t kid[8];
    double val[8];
    uint32_t hk[8];
    int nk = 0;
    uint32_t rest = m;
    while (rest && nk < 8) {
        const uint32_t b = 31u - (uint32_t)__clz((int)rest);
        rest ^= 1u << b;
        kid[nk++] = f.col[s0 + b];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        val[k] = 0.0; hk[k] = 1u;
        if (k < nk && !(has_chain && kid[k] == p + 1u)) {
            const uint32_t c = kid[k];
            val[k] = fl_ld_cg(&f.A[c + dep]);
            hk[k] = fl_ld_cg(&f.hgt[c + dep]) + 1u;
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k < nk) {
            if (has_chain && kid[k] == p + 1u) { seen = true; }
            else {
                if (hk[k] > hmax) hmax = hk[k];
                if (!seen) pre += val[k];
                else {
                    if (np == 0) p1 = val[k]; else if (np == 1) p2 = val[k]; else if (np < FL_MAXPOST) xp[np - 2] = val[k];
                    ++np;
                }
            }
        }
    }
    if (rest) np = 255u;  // more than 8 children: slow path
    while (rest) {  // more than 8 children: one at a time
        const uint32_t b = 31u - (uint32_t)__clz((int)rest);
        rest ^= 1u << b;
        const uint32_t c = f.col[s0 + b];
        if (has_chain && c == p + 1u) { seen = true; continue; }
        const double v = fl_ld_cg(&f.A[c + dep]);
        const uint32_t hc = fl_ld_cg(&f.hgt[c + dep]) + 1u;
        if (hc > hmax) hmax = hc;
        if (!seen) pre += v;
    }
    return np > FL_MAXPOST ? 15u : np;
}

// np == 15: add every child after the chain child, in order, straight from the children
__device__ __noinline__ double fl_add_posts_slow(const FlFlow& f, uint32_t p, double y) {
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

// release fence at gpu scope (lighter than __threadfence(), which is fence.sc)
__device__ __forceinline__ void fl_fence_release() {
#ifndef FL_EMU
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
}

// gather the partial sums of p and store them; returns the state word that publishes them
__device__ __forceinline__ uint32_t fl_gather_store(const FlFlow& f, uint32_t p, bool has_chain, uint32_t dep) {
    double pre, p1, p2;
    uint32_t hp;
    const uint32_t np = fl_gather_lights(f, p, has_chain, dep, pre, p1, p2, hp);
    f.pre[p] = pre;
    f.hpre[p] = hp;
    if (np >= 1u && np != 15u) f.post1[p] = p1;
    if (np >= 2u && np != 15u) f.post2[p] = p2;
    return fl_st_publish(np, hp);
}

// A finished segment head `h` (area y already stored) reports to its receiver p.  Returns the site where the
// next climb starts (the caller continues there), or FL_NONE when this flow ends.  *dep_out carries the
// opaque zero that orders the next climb's loads after the deciding atomic.  `fenced`: the caller has already
// fenced its stores.  The static data of p (counts, row, child mask) is requested right behind the atomic so
// that the round trips overlap.
// Ordering (FL_ATOMIC_ORDERING, the default): the two deciding atomics are themselves the release / acquire
// operations -- atom.acq_rel.gpu = MEMBAR.ALL.GPU ; ATOMG ; CCTL.IVALL in SASS, i.e. ONE memory barrier per atomic
// (release side) and an L1 invalidation (acquire side).  The fence form, fence.acq_rel.gpu before and after a relaxed
// atomic, is the same thing to the memory model but costs two barriers per atomic: ncu put 25 % of the warp kernel's
// stall samples on those MEMBARs (profiles/r2i_src_k_area_flow_long.txt).  -DFL_ATOMIC_ORDERING=0 builds the fence form.
#ifndef FL_ATOMIC_ORDERING
#define FL_ATOMIC_ORDERING 1
#endif
#if defined(FL_EMU) || !FL_ATOMIC_ORDERING || FL_RELAXED_READERS
#define FL_QUALIFIED_ATOMICS 0
#else
#define FL_QUALIFIED_ATOMICS 1
__device__ __forceinline__ uint32_t fl_atom_add_acq_rel(uint32_t* p, uint32_t v) {
    uint32_t o;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory");
    return o;
}
__device__ __forceinline__ uint32_t fl_atom_add_acquire(uint32_t* p, uint32_t v) {
    uint32_t o;
    asm volatile("atom.acquire.gpu.global.add.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory");
    return o;
}
#endif
__device__ __forceinline__ uint32_t fl_report(const FlFlow& f, uint32_t h, uint32_t p, bool fenced, uint32_t* dep_out) {
    (void)h;
#if FL_QUALIFIED_ATOMICS
    // release: publishes A[h], hgt[h] (and, by cumulativity, what the caller's warp stored before its __syncwarp());
    // acquire: the last reporter sees the other reporters' A / hgt
    const uint32_t prev = fenced ? fl_atom_add_acquire(&f.state[p], 1u) : fl_atom_add_acq_rel(&f.state[p], 1u);
#else
    if (!fenced) fl_fence_release();  // publish A[h], hgt[h] (and, by cumulativity, what the caller's warp stored
                                      // before its __syncwarp())
    const uint32_t prev = atomicAdd(&f.state[p], 1u);
#endif
    const uint32_t nw = f.nwait[p] & FL_NW_COUNT;
    const uint32_t sh = f.seg_head[p];
    const bool p_has_chain = (p + 1u < f.n) && (f.recv[p + 1u] == p);
    if ((prev & FL_ST_COUNT_MASK) + 1u < nw) return FL_NONE;
    // last reporter at p: gather and publish p's partial sums
    const uint32_t swait = f.seg_wait[sh];
#if FL_QUALIFIED_ATOMICS
    f.state[p] = fl_gather_store(f, p, p_has_chain, 0u);  // no report can follow the last one: a plain store
    // release: the partial sums are published before the segment counter moves; acquire: the last publisher sees the
    // other publishers' partial sums
    const uint32_t done = fl_atom_add_acq_rel(&f.seg_done[sh], 1u) + 1u;
    if (done < swait) return FL_NONE;
#else
    fl_fence_acquire();  // the other reporters' A / hgt: acquire side of their fence + atomic on state[p]
    f.state[p] = fl_gather_store(f, p, p_has_chain, fl_dep0(prev));  // no report can follow the last one: a plain store
    fl_fence_release();  // publish the partial sums before the segment counter moves
    const uint32_t done = atomicAdd(&f.seg_done[sh], 1u) + 1u;
    if (done < swait) return FL_NONE;
    fl_fence_acquire();  // the other publishers' partial sums: acquire side of their fence + atomic on seg_done[sh]
#endif
    *dep_out = fl_dep0(done);
    return fl_seg_start(f, sh);  // every waiting site of the segment is published: climb it
}

// ------------------------------------------------------------------------------------------------
// thread-level flow: climb the segment whose tail is `cur`, report, possibly climb the next segment, ...
// ------------------------------------------------------------------------------------------------
__device__ void fl_flow_thread(const FlFlow& f, uint32_t cur, double x, uint32_t hrun, bool has_chain, bool may_park) {
    uint32_t climbed = 0;
    if (may_park) FL_TLOG(f, cur, 0);
    for (;;) {
        double y = 0.0;
        uint32_t p = FL_NONE;
        bool at_head = false;
        // ---- up to FL_TB consecutive sites of the chain, loads issued up front ----
        const uint32_t nb = cur + 1u < (uint32_t)FL_TB ? cur + 1u : (uint32_t)FL_TB;
        uint32_t cm[FL_TB], rc[FL_TB], st[FL_TB];
        double ar[FL_TB], pr[FL_TB], q1[FL_TB], q2[FL_TB];
#pragma unroll
        for (int k = 0; k < FL_TB; ++k) {
            cm[k] = 0u; rc[k] = FL_NONE; ar[k] = 0.0; st[k] = 0u; pr[k] = 0.0; q1[k] = 0.0; q2[k] = 0.0;
            if ((uint32_t)k < nb) {
                const uint32_t i = cur - k;
                cm[k] = f.cmask[i]; rc[k] = f.recv[i]; ar[k] = f.areas[i];
                st[k] = fl_ld_cg(&f.state[i]); pr[k] = fl_ld_cg(&f.pre[i]); q1[k] = fl_ld_cg(&f.post1[i]);
            }
        }
        uint32_t done = 0;
        FL_COUNT(f, FLS_T_BATCH, 1);
#pragma unroll
        for (int k = 0; k < FL_TB; ++k) {
            if (!at_head && done == (uint32_t)k && (uint32_t)k < nb) {
                const uint32_t idx = cur - k;
                const bool hc = (k > 0) || has_chain;
                const uint32_t nl = (uint32_t)__popc(cm[k]) - (hc ? 1u : 0u);
                if (nl == 0u) {
                    y = hc ? (ar[k] + x) : ar[k];
                } else {
                    if (!(st[k] & FL_ST_PRE_READY)) {  // cannot happen (a segment is climbed when complete)
                        FL_COUNT(f, FLS_T_NOTREADY, 1);
                        atomicOr(&f.flags[FL_FLAG_BROKEN], 1u);
                        return;
                    }
                    const uint32_t npk = fl_st_np(st[k]);
                    y = hc ? (pr[k] + x) : pr[k];
                    if (npk == 15u) y = fl_add_posts_slow(f, idx, y);
                    else if (npk >= 1u) {
                        y += q1[k];
                        if (npk >= 2u) {
                            const double p2v = fl_ld_cg(&f.post2[idx]);
                            double xv[FL_XPOST];
#pragma unroll
                            for (int j = 0; j < FL_XPOST; ++j)
                                xv[j] = ((uint32_t)j + 2u < npk) ? fl_ld_cg(&f.xpost[(size_t)idx * FL_XPOST + j]) : 0.0;
                            y += p2v;
#pragma unroll
                            for (int j = 0; j < FL_XPOST; ++j)
                                if ((uint32_t)j + 2u < npk) y += xv[j];
                        }
                    }
                    const uint32_t hq = fl_st_hp(st[k], f.hpre, idx);
                    if (hq > hrun) hrun = hq;
                }
                f.A[idx] = y;
                f.hsuf[idx] = hrun;
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
        FL_COUNT(f, FLS_T_SITES, done + (at_head ? 1u : 0u));
        if (done > 0u) has_chain = true;
        cur -= done;
        if (!at_head) {  // whole batch climbed
            climbed += done;
            if (may_park && f.park_after != 0u && climbed >= f.park_after) {
                FL_COUNT(f, FLS_T_PARKED, 1);
                FL_TLOG(f, cur, 1);
                f.xbuf[cur] = x;  // a long chain: leave the rest to the warp-level pass
                f.hbuf[cur] = hrun;
                const uint32_t slot = atomicAdd(&f.counters[0], 1u);
                f.parked[slot] = cur;
                return;
            }
            continue;
        }
        // ---- `cur` is the segment head with final area y; p = recv[cur] ----
        FL_COUNT(f, FLS_T_HEADS, 1);
        FL_TLOG(f, cur, 3);
        f.hgt[cur] = hrun;
        if (p == cur) {  // tree root: its segment has the largest nesting height of the tree
            if (hrun > 0u) atomicMax(&f.flags[FL_FLAG_MAXDEPTH], hrun);
            return;
        }
        uint32_t dep = 0u;
        const uint32_t next_tail = fl_report(f, cur, p, false, &dep);
        if (next_tail == FL_NONE) return;
        FL_COUNT(f, FLS_T_SEGSTART, 1);
        cur = next_tail + dep; climbed = 0u;
        FL_TLOG(f, cur, 0);
        fl_seg_entry(f, cur, x, hrun, has_chain);
    }
}

// pre-pass A: how many children will REPORT to each site = its non-leaf segment heads.
// (Leaves are read by their parent directly; chain children hand over inside the segment.)
__global__ void __launch_bounds__(256) k_count_waits(uint32_t n, const uint32_t* __restrict__ recv,
                                                      const uint32_t* __restrict__ cmask,
                                                      const double* __restrict__ areas, uint32_t* nwait,
                                                      double* __restrict__ A, uint32_t* __restrict__ hgt,
                                                      uint32_t* __restrict__ hsuf) {
    const uint32_t q = FL_TID;
    if (q >= n) return;
    const uint32_t p = recv[q];
    const bool chained = q > 0u && p == q - 1u;
    if (cmask[q] == 0u) {
        if (!chained) { A[q] = areas[q]; hgt[q] = 0u; hsuf[q] = 0u; }  // a leaf that is a segment of its own
        return;
    }
    if (p == q || chained) return;  // root, or chained to its receiver
    atomicAdd(&nwait[p], 1u);
}

// pre-pass B: scan input for the segment heads: a site that is not chained to its receiver starts a segment
__global__ void __launch_bounds__(256) k_seg_keys(uint32_t n, const uint32_t* __restrict__ recv,
                                                   uint32_t* __restrict__ key) {
    const uint32_t q = FL_TID;
    if (q >= n) return;
    key[q] = (q > 0u && recv[q] == q - 1u) ? 0u : q;  // inclusive max-scan of this = head of q's segment
}

// pre-pass C: per segment (stored at its head): the tail, and the number of waiting sites.  Sites whose non-chain
// children are all leaves are published right here (they wait for nobody).
__global__ void __launch_bounds__(256) k_seg_prepare(FlFlow f, uint32_t* seg_tail, uint32_t* seg_wait) {
    const uint32_t q = FL_TID;
    if (q >= f.n) return;
    const uint32_t cm = f.cmask[q];
    const bool has_chain = (q + 1u < f.n) && (f.recv[q + 1u] == q);
    const uint32_t sh = f.seg_head[q];
    if (!has_chain) seg_tail[sh] = q;
    if (cm == 0u) return;
    if (f.nwait[q] != 0u) { atomicAdd(&seg_wait[sh], 1u); return; }
    if ((uint32_t)__popc(cm) - (has_chain ? 1u : 0u) == 0u) return;  // only the chain child
    f.state[q] = fl_gather_store(f, q, has_chain, 0u);
}

// pass 1: one thread per segment TAIL whose segment waits for nobody; all other segments are started by the
// thread that publishes their last waiting site.  One-site leaf segments only publish their own area / height.
__global__ void __launch_bounds__(256) k_area_flow(FlFlow f) {
    const uint32_t q = FL_TID;
    if (q >= f.n) return;
    if ((q + 1u < f.n) && (f.recv[q + 1u] == q)) return;  // not a tail
    const uint32_t sh = f.seg_head[q];
    if (sh == q && f.cmask[q] == 0u) return;  // a leaf that is a segment of its own: done in k_count_waits
    if (f.seg_wait[sh] != 0u) return;
    FL_COUNT(f, FLS_T_CLIMBS, 1);
    fl_flow_thread(f, q, 0.0, 0u, false, true);
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

// ------------------------------------------------------------------------------------------------
// warp-level flow.  All lanes hold identical copies of the flow state (cur, x, hrun, ...).
// ------------------------------------------------------------------------------------------------
struct FlWin {  // one 32-site window of a chain: lane l holds site base - l
    uint32_t cm, rc, st;
    double ar, pre, p1, p2;
};

__device__ __forceinline__ FlWin fl_win_load(const FlFlow& f, uint32_t base, int lane) {
    FlWin w;
    w.cm = 0u; w.rc = FL_NONE; w.st = 0u; w.ar = 0.0; w.pre = 0.0; w.p1 = 0.0; w.p2 = 0.0;
    const long long li = (long long)base - lane;
    if (li >= 0) {
        const uint32_t idx = (uint32_t)li;
        w.cm = f.cmask[idx];
        w.rc = f.recv[idx];
        w.ar = f.areas[idx];
        w.st = fl_ld_cg(&f.state[idx]);
        w.pre = fl_ld_cg(&f.pre[idx]);
        w.p1 = fl_ld_cg(&f.post1[idx]);
        w.p2 = fl_ld_cg(&f.post2[idx]);
    }
    return w;
}

#define FL_WDEPTH 3  // windows kept in flight on a long chain
#ifndef FL_XPOST_PREFETCH
#define FL_XPOST_PREFETCH 0  // build-time A/B switch (tools/ab_build.py): xpost of the next window requested one window
                             // ahead -- measured SLOWER (K4 0.319 vs 0.300 ms per iteration at 1M sites, profiles/r2j_ab.txt)
#endif

__device__ void fl_flow_warp(const FlFlow& f, uint32_t cur, double x, uint32_t hrun, bool has_chain, FlAreaSmem& sm) {
    const int lane = threadIdx.x & 31;
    // ring[j] holds the window whose lane 0 is site cur - 32*j, for j < nring
    FlWin ring[FL_WDEPTH];
    uint32_t nring = 0;
#pragma unroll
    for (int j = 0; j < FL_WDEPTH; ++j) { ring[j].cm = 0u; ring[j].rc = FL_NONE; ring[j].st = 0u; ring[j].ar = 0.0; ring[j].pre = 0.0; ring[j].p1 = 0.0; ring[j].p2 = 0.0; }
    double xnext[FL_XPOST];  // xpost of the window after the current one (valid while the chain goes on)
    bool xnext_valid = false;
#pragma unroll
    for (int j = 0; j < FL_XPOST; ++j) xnext[j] = 0.0;
    for (;;) {
        const long long t_win = FL_CLOCK();
        const bool first_win = nring == 0u;
        if (nring == 0u) { ring[0] = fl_win_load(f, cur, lane); nring = 1u; }
        const FlWin win = ring[0];
        const long long li = (long long)cur - lane;
        const bool valid = li >= 0;
        const uint32_t idx = (uint32_t)li;
        const uint32_t rc_up = __shfl_up_sync(FL_FULL, win.rc, 1);
        const bool link = valid && (lane == 0 || rc_up == idx);
        const uint32_t linkmask = __ballot_sync(FL_FULL, link);
        const uint32_t nproc = linkmask == FL_FULL ? 32u : (uint32_t)__ffs((int)~linkmask) - 1u;  // sites of the chain
        const bool hc = (lane > 0) || has_chain;
        const bool inwin = (uint32_t)lane < nproc;
        const uint32_t nl = inwin ? (uint32_t)__popc(win.cm) - (hc ? 1u : 0u) : 0u;
        const bool lit = inwin && nl > 0u;
        if (__ballot_sync(FL_FULL, lit && !(win.st & FL_ST_PRE_READY))) {  // cannot happen
            if (lane == 0) { FL_COUNT(f, FLS_W_NOTREADY, 1); atomicOr(&f.flags[FL_FLAG_BROKEN], 1u); }
            return;
        }
        const bool climbs = valid && idx > 0u && win.rc == idx - 1u;
        // keep FL_WDEPTH windows in flight while the chain goes on
        const bool goes_on = nproc == 32u && __shfl_sync(FL_FULL, (int)climbs, 31) && cur >= 32u;
        if (goes_on) {
#pragma unroll
            for (int j = 1; j < FL_WDEPTH; ++j)
                if (nring == (uint32_t)j && cur >= 32u * (uint32_t)j) { ring[j] = fl_win_load(f, cur - 32u * (uint32_t)j, lane); nring = (uint32_t)j + 1u; }
        }
        double b = win.ar, q1 = 0.0, q2 = 0.0;
        uint32_t npk = 0u, hq = 0u;
        if (lit) {
            npk = fl_st_np(win.st);
            b = win.pre;
            if (npk >= 1u && npk != 15u) q1 = win.p1;
            if (npk >= 2u && npk != 15u) q2 = win.p2;
            hq = (win.st & FL_ST_HP_MASK) >> FL_ST_HP_SHIFT;
            if (hq == FL_ST_HP_OVER) hq = fl_ld_cg(&f.hpre[idx]);
        }
        if (lane == 0) { FL_COUNT(f, FLS_W_WINDOWS, 1); FL_COUNT(f, FLS_W_SITES, nproc); }
        // The first site of a segment without a chain child starts from x = 0.0: 0.0 + b == b exactly (areas are
        // positive), so the chain below needs no special case for it.
        if (!has_chain) x = 0.0;
        // my site's terms: b, then the children after the chain child
        uint32_t nterm = 0u;
        bool slow = false;
        double xv[FL_XPOST];
#pragma unroll
        for (int j = 0; j < FL_XPOST; ++j) xv[j] = xnext[j];  // (requested while the previous window was climbed)
        if (inwin) {
            nterm = 1u;
            if (lit) {
                if (npk == 15u) slow = true;
                else {
                    nterm += npk;
                    if (npk > 2u && !xnext_valid) {
#pragma unroll
                        for (int j = 0; j < FL_XPOST; ++j)
                            if ((uint32_t)j + 2u < npk) xv[j] = fl_ld_cg(&f.xpost[(size_t)idx * FL_XPOST + j]);
                    }
                }
            }
        }
        // The 3rd.. children of the NEXT window's sites: requested now, used one window later.  Every site of a segment
        // is published before its climb starts, so what the ring holds of the next window is final.
        xnext_valid = false;
        if (FL_XPOST_PREFETCH && goes_on && nring >= 2u) {
            xnext_valid = true;
            const long long li1 = (long long)cur - 32 - lane;
            const uint32_t st1 = ring[1].st, np1 = fl_st_np(st1);
#pragma unroll
            for (int j = 0; j < FL_XPOST; ++j) {
                xnext[j] = 0.0;
                if (li1 >= 0 && (st1 & FL_ST_PRE_READY) && np1 != 15u && (uint32_t)j + 2u < np1)
                    xnext[j] = fl_ld_cg(&f.xpost[(size_t)li1 * FL_XPOST + j]);
            }
        }
        const uint32_t slowmask = __ballot_sync(FL_FULL, slow);
        uint32_t off = nterm;  // inclusive scan over the lanes -> where my terms go
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t w = __shfl_up_sync(FL_FULL, off, o);
            if (lane >= o) off += w;
        }
        const uint32_t nall = __shfl_sync(FL_FULL, off, 31);
        off -= nterm;
        double mine = 0.0;
        if (slowmask == 0u) {
            __syncwarp();
            if (inwin) {
                sm.t[off] = b;
                if (lit) {
                    if (npk >= 1u) sm.t[off + 1u] = q1;
                    if (npk >= 2u) sm.t[off + 2u] = q2;
#pragma unroll
                    for (int j = 0; j < FL_XPOST; ++j)
                        if ((uint32_t)j + 2u < npk) sm.t[off + 3u + (uint32_t)j] = xv[j];
                }
            }
            if (lane < 8) sm.t[nall + (uint32_t)lane] = 0.0;  // padding: r + 0.0 == r
            __syncwarp();
            // the serial chain, identically in every lane: 8 terms fetched together, then 8 dependent additions
            double r = x;
            for (uint32_t t0 = 0; t0 < nall; t0 += 8u) {
                double v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = sm.t[t0 + (uint32_t)j];
#pragma unroll
                for (int j = 0; j < 8; ++j) { r = r + v[j]; v[j] = r; }
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) sm.p[t0 + (uint32_t)j] = v[j];
                }
            }
            x = r;
            __syncwarp();
            if (inwin) mine = sm.p[off + nterm - 1u];
        } else {
            // a site with more than FL_MAXPOST children after its chain child (or more than 8 children): rare
            double r = x;
            for (uint32_t k = 0; k < nproc; ++k) {
                r = fl_shfl(b, (int)k) + r;
                const uint32_t npk_k = __shfl_sync(FL_FULL, lit ? npk : 0u, (int)k);
                if (npk_k >= 3u) {  // all children after the chain child, straight from the children
                    double v = 0.0;
                    if ((uint32_t)lane == k) v = fl_add_posts_slow(f, idx, r);
                    __syncwarp();
                    r = fl_shfl(v, (int)k);
                } else {
                    const double a1 = fl_shfl(q1, (int)k), a2 = fl_shfl(q2, (int)k);
                    if (npk_k >= 1u) r += a1;
                    if (npk_k >= 2u) r += a2;
                }
                if ((uint32_t)lane == k) mine = r;
            }
            x = r;
        }
        // running nesting height after every site of the window (lane 0 is climbed first)
        uint32_t hs = inwin ? hq : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t w = __shfl_up_sync(FL_FULL, hs, o);
            if (lane >= o && w > hs) hs = w;
        }
        if (hrun > hs) hs = hrun;
        hrun = __shfl_sync(FL_FULL, hs, 31);
        if (inwin) {
            f.A[idx] = mine;
            f.hsuf[idx] = hs;
            if (climbs) f.hgt[idx] = FL_NONE;
        }
        if (lane == 0) FL_COUNT(f, first_win ? FLS_W_CYC_FIRSTWIN : FLS_W_CYC_WIN, FL_CLOCK() - t_win);
        const int lastl = (int)nproc - 1;  // nproc >= 1: lane 0 is always on the chain
        const int last_climbs = __shfl_sync(FL_FULL, (int)climbs, lastl);
        if (last_climbs) {  // nproc == 32: whole window climbed, shift the ring
            has_chain = true;
            cur -= 32u;
#pragma unroll
            for (int j = 0; j + 1 < FL_WDEPTH; ++j) ring[j] = ring[j + 1];
            nring = nring > 0u ? nring - 1u : 0u;
            continue;
        }
        // lane `lastl` is the segment head
        const uint32_t h = cur - (uint32_t)lastl;
        const uint32_t p = __shfl_sync(FL_FULL, win.rc, lastl);
        nring = 0u;
        if (lane == 0) { f.hgt[h] = hrun; FL_COUNT(f, FLS_W_HEADS, 1); FL_TLOG(f, h, 3); }
        if (p == h) {
            if (lane == 0 && hrun > 0u) atomicMax(&f.flags[FL_FLAG_MAXDEPTH], hrun);
            return;
        }
        uint32_t next_tail = FL_NONE, dep = 0u;
        const long long t_rep = FL_CLOCK();
        __syncwarp();  // the lanes' stores (A, hsuf, hgt) happen before lane 0's release fence inside fl_report
        if (lane == 0) next_tail = fl_report(f, h, p, false, &dep);
        next_tail = __shfl_sync(FL_FULL, next_tail, 0);
        __syncwarp();  // lane 0's acquire fence happens before the other lanes' loads of the next segment
        if (lane == 0) FL_COUNT(f, FLS_W_CYC_REPORT, FL_CLOCK() - t_rep);
        if (next_tail == FL_NONE) return;
        if (lane == 0) FL_COUNT(f, FLS_W_SEGSTART, 1);
        cur = next_tail + fl_dep0(next_tail);  // the broadcast value came after lane 0's atomics
        if (lane == 0) FL_TLOG(f, cur, 0);
        fl_seg_entry(f, cur, x, hrun, has_chain);
    }
}
#endif

// pass 2: the parked (long) climbs, continued by whole warps (emulation: threads); every warp takes the next
// parked climb through one atomic counter until none is left.
__global__ void __launch_bounds__(256, 2) k_area_flow_long(FlFlow f) {
    const uint32_t count = fl_ld_cg(&f.counters[0]);
#ifdef FL_EMU
    for (uint32_t i = FL_TID; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t cur = f.parked[i];
        fl_flow_thread(f, cur, f.xbuf[cur], f.hbuf[cur], true, false);
    }
#else
    __shared__ FlAreaSmem chain_smem[8];  // one per warp (256 threads)
    FlAreaSmem& sm = chain_smem[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    for (;;) {
        uint32_t i = 0u;
        if (lane == 0) i = atomicAdd(&f.counters[1], 1u);
        i = __shfl_sync(FL_FULL, i, 0);
        if (i >= count) return;
        const uint32_t cur = f.parked[i];
        if (lane == 0) { FL_COUNT(f, FLS_W_FLOWS, 1); FL_TLOG(f, cur, 2); }
        const long long t_flow = FL_CLOCK();
        fl_flow_warp(f, cur, fl_ld_cg(&f.xbuf[cur]), fl_ld_cg(&f.hbuf[cur]), true, sm);
        if (lane == 0) FL_COUNT(f, FLS_W_CYC_FLOW, FL_CLOCK() - t_flow);
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// Incremental K4.  The drainage area of a site is a function of its subtree only, so between two
// iterations it changes only at the ancestors of sites whose receiver changed (a few percent of the sites
// after the first iterations).  Everything the full pass leaves behind -- A, the published partial sums,
// hgt at the heads, hsuf along the chains -- stays valid elsewhere, in the same numbering.  What is redone:
//   * regather list: the old and the new receiver of every re-routed site (their child sets changed) and the
//     receiver of every dirty segment head (a child's area / height changed);
//   * per segment, the sites from the highest regathered one down to the head are climbed again, starting
//     from the stored area / running height of the clean chain above (fl_seg_entry);
//   * the dataflow between dirty segments is the one of the full pass: a dirty head reports to its receiver,
//     the last reporter regathers, a segment is climbed once when its waiting sites are all published.
// Every addition is still the reference's, in the reference's order; clean sites simply keep a result that a
// recomputation would reproduce bit for bit.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fl_mark_up(const FlFlow& f, uint32_t p) {
    for (;;) {
        // two dependent round trips per step up the forest: {flag, segment head} then {dirty mark, the head's receiver};
        // the results are looked at only after everything has been requested
        const uint32_t old = atomicOr(&f.nwait[p], FL_NW_RFLAG);
        const uint32_t sh = f.seg_head[p];
        const uint32_t prev = atomicMax(&f.dirty_from[sh], p + 1u);
        const uint32_t pp = f.recv[sh];
        if (!(old & FL_NW_RFLAG)) f.rlist[atomicAdd(&f.flags[FL_FLAG_NREGATHER], 1u)] = p;
        if (prev != 0u) return;  // the segment is already dirty: whoever marked it first walks on from its head
        f.slist[atomicAdd(&f.flags[FL_FLAG_NDIRTY], 1u)] = sh;
        if (pp == sh) return;  // tree root
        atomicAdd(&f.nwait[pp], 1u);  // the dirty head sh will report to pp
        p = pp;
    }
}

// one thread per re-routed site (list written by K1)
__global__ void __launch_bounds__(128) k_incr_mark(FlFlow f, uint32_t nchg, const uint32_t* __restrict__ chg_node,
                                                    const uint32_t* __restrict__ chg_old) {
    const uint32_t k = FL_TID;
    if (k >= nchg) return;
    const uint32_t q = chg_node[k], po = chg_old[k], pn = f.recv[q];
    if (q > 0u && pn == q - 1u) f.hgt[q] = FL_NONE;         // now chained to its receiver: no longer a head
    else if (q > 0u && po == q - 1u) f.hgt[q] = f.hsuf[q];  // the chain broke below q: q heads what is left of it
    if (pn != q) fl_mark_up(f, pn);
    if (po != q) fl_mark_up(f, po);
}

// regather list: sites that wait for dirty heads are counted on their segment, the others are published here
__global__ void __launch_bounds__(128) k_incr_prepare(FlFlow f) {
    const uint32_t cnt = f.flags[FL_FLAG_NREGATHER];
    for (uint32_t k = FL_TID; k < cnt; k += gridDim.x * blockDim.x) {
        const uint32_t p = f.rlist[k];
        f.state[p] = 0u;
        if (f.nwait[p] & FL_NW_COUNT) { atomicAdd(&f.seg_wait[f.seg_head[p]], 1u); continue; }
        const uint32_t cm = f.cmask[p];
        const bool has_chain = (p + 1u < f.n) && (f.recv[p + 1u] == p);
        if (cm == 0u || (uint32_t)__popc(cm) - (has_chain ? 1u : 0u) == 0u) continue;  // nothing to gather
        f.state[p] = fl_gather_store(f, p, has_chain, 0u);
    }
}

// dirty segments that wait for nobody start right away; the others are started by their last publisher
__global__ void __launch_bounds__(64) k_incr_start(FlFlow f) {
    const uint32_t cnt = f.flags[FL_FLAG_NDIRTY];
    for (uint32_t k = FL_TID; k < cnt; k += gridDim.x * blockDim.x) {
        const uint32_t sh = f.slist[k];
        if (f.seg_wait[sh] != 0u) continue;
        const uint32_t cur = fl_seg_start(f, sh);
        double x;
        uint32_t hrun;
        bool has_chain;
        fl_seg_entry(f, cur, x, hrun, has_chain);
        FL_COUNT(f, FLS_T_CLIMBS, 1);
        fl_flow_thread(f, cur, x, hrun, has_chain, true);
    }
}

// leave the bookkeeping zeroed for the next incremental pass
__global__ void __launch_bounds__(256) k_incr_cleanup(FlFlow f) {
    const uint32_t nr = f.flags[FL_FLAG_NREGATHER], ns = f.flags[FL_FLAG_NDIRTY];
    for (uint32_t k = FL_TID; k < nr; k += gridDim.x * blockDim.x) f.nwait[f.rlist[k]] = 0u;
    for (uint32_t k = FL_TID; k < ns; k += gridDim.x * blockDim.x) {
        const uint32_t sh = f.slist[k];
        f.dirty_from[sh] = 0u;
        f.seg_wait[sh] = 0u;
        f.seg_done[sh] = 0u;
    }
}

// ------------------------------------------------------------------------------------------------
// First iteration: there are no previous drainage areas to choose heavy children by, so the layout is built from
// SUBTREE SIZES of the first forest -- integer counts, order-free, one counting sweep from the leaves (the last child
// to arrive carries the sum on).  Any choice of heavy children is correct; sizes keep the nesting logarithmic, which
// lets the first iteration (the deepest forest of the run) use the dataflow sweeps instead of one launch per level.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_subtree_init(uint32_t n, const uint32_t* __restrict__ cmask,
                                                       uint32_t* __restrict__ size, uint32_t* __restrict__ pend) {
    const uint32_t v = FL_TID;
    if (v >= n) return;
    size[v] = 1u;
    const uint32_t k = (uint32_t)__popc(cmask[v]);
    pend[v] = k ? k + 1u : 0u;  // biased by one: 0 stays "leaf" while the sweep runs
}
__global__ void __launch_bounds__(256) k_subtree_sweep(uint32_t n, const uint32_t* __restrict__ recv,
                                                        const uint32_t* __restrict__ cmask, uint32_t* size,
                                                        uint32_t* pend) {
    uint32_t v = FL_TID;
    if (v >= n) return;
    if (cmask[v] != 0u) return;  // not a leaf
    uint32_t sz = 1u;
    for (;;) {
        const uint32_t p = recv[v];
        if (p == v) return;
        atomicAdd(&size[p], sz);
        __threadfence();
        if (atomicSub(&pend[p], 1u) != 2u) return;  // somebody else arrives last
        __threadfence();  // acquire side: the other children's additions to size[p]
        sz = atomicAdd(&size[p], 0u);
        v = p;
    }
}
__global__ void __launch_bounds__(256) k_subtree_weight(uint32_t n, const uint32_t* __restrict__ size,
                                                         double* __restrict__ weight) {
    const uint32_t v = FL_TID;
    if (v < n) weight[v] = (double)size[v];
}

// ------------------------------------------------------------------------------------------------
// K5 on dynamic segments (generator.rs:162-203)
// ------------------------------------------------------------------------------------------------
struct FlElev {
    uint32_t n;
    const uint32_t* recv;
    const double* drecv;
    const double* tcel;  // 1/(k*sqrt(A))*d per site (k_celerity_term)
    const double* uplift;
    const double* tan_slope;  // may be null
    const uint8_t* is_outlet;
    double* elev;
    double* rt;
    uint32_t* root_of;
    uint32_t* flags;
    uint32_t* lvl;       // out: nesting height of the segment each site belongs to (schedules the next K4)
    uint32_t lvl_value;  // the height this launch works on
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

// B sites of one segment by one thread; returns true when the segment ended inside the batch
template <int B>
__device__ __forceinline__ bool fl_elev_batch(const FlElev& e, uint32_t& q, uint32_t h, bool is_root, uint32_t root,
                                              double& rt_prev, double& z_prev, double& e_out, double& rt_out,
                                              bool& changed) {
    const uint32_t nb = e.n - q < (uint32_t)B ? e.n - q : (uint32_t)B;
    double d[B], t[B], up[B], eo[B], ms[B];
    uint32_t nx[B];
#pragma unroll
    for (int k = 0; k < B; ++k) {
        d[k] = 1.0; t[k] = 0.0; up[k] = 0.0; eo[k] = 0.0; ms[k] = 0.0; nx[k] = FL_NONE;
        if ((uint32_t)k < nb) {
            const uint32_t i = q + k;
            t[k] = e.tcel[i];
            up[k] = e.uplift[i];
            eo[k] = e.elev[i];
            if (e.tan_slope) { ms[k] = e.tan_slope[i]; d[k] = e.drecv[i]; }
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
            if (is_root && i == h) e_out = z;  // later sites read elevations[outlet] after the outlet's own update
            e.elev[i] = z;
            e.rt[i] = rti;
            e.root_of[i] = root;
            e.lvl[i] = e.lvl_value;
            rt_prev = rti;
            z_prev = z;
            if (nx[k] != i) ended = true;
        }
    }
    q += nb;
    return ended || q >= e.n;
}

#ifndef FL_EMU
struct FlEWin {  // one 32-site window of a segment: lane l holds site base + l
    double t, up, eold, ms, d;
    uint32_t nx;
    bool valid;
};

__device__ __forceinline__ FlEWin fl_ewin_load(const FlElev& e, uint32_t base, int lane) {
    FlEWin w;
    w.t = 0.0; w.up = 0.0; w.eold = 0.0; w.ms = 0.0; w.d = 1.0; w.nx = FL_NONE;
    const unsigned long long i64 = (unsigned long long)base + (unsigned)lane;
    w.valid = i64 < e.n;
    if (w.valid) {
        const uint32_t i = (uint32_t)i64;
        w.t = e.tcel[i];
        w.up = e.uplift[i];
        w.eold = e.elev[i];
        w.nx = (i + 1u < e.n) ? e.recv[i + 1u] : FL_NONE;
        if (e.tan_slope) { w.ms = e.tan_slope[i]; w.d = e.drecv[i]; }
    }
    return w;
}

// one window of a long segment: the two serial chains (response time; clamp if max_slope).  The per-lane
// terms are staged in shared memory and every lane runs the identical chain over broadcast reads, so that
// the dependent double additions are the only thing on the critical path (no shuffles, no selects).

__device__ __forceinline__ void fl_elev_window(const FlElev& e, const FlEWin& w, uint32_t q, uint32_t nproc, int lane,
                                               uint32_t root, double& rt_prev, double& z_prev, double e_out,
                                               double rt_out, bool& changed, FlChainSmem& sm) {
    __syncwarp();
    sm.in[lane] = ((uint32_t)lane < nproc) ? w.t : 0.0;  // padding: 0.0 + (r + 0.0) == r (r >= +0.0)
    __syncwarp();
    {
        // 8 terms fetched together, then the dependent additions (identically in every lane)
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
    }
    if ((uint32_t)lane < nproc) {
        const uint32_t i = q + (uint32_t)lane;
        changed |= (z != w.eold);
        e.elev[i] = z;
        e.rt[i] = my_rt;
        e.root_of[i] = root;
        e.lvl[i] = e.lvl_value;
    }
}

__device__ __forceinline__ uint32_t fl_elev_nproc(const FlEWin& w, uint32_t q, int lane, uint32_t& endmask) {
    const uint32_t i = q + (uint32_t)lane;
    endmask = __ballot_sync(FL_FULL, !w.valid || w.nx != i);
    if (!endmask) return 32u;
    const int el = __ffs((int)endmask) - 1;
    const int el_valid = __shfl_sync(FL_FULL, (int)w.valid, el);
    return (uint32_t)el + (el_valid ? 1u : 0u);
}

// rest of a long segment, walked by the whole warp.  Returns "changed".  The serial chains of a window are
// much shorter than a DRAM round trip, so windows are fetched FL_EDEPTH ahead (registers) once the segment
// has proved to be longer than one window.
#define FL_EDEPTH 4
__device__ bool fl_elev_warp(const FlElev& e, uint32_t q, uint32_t root, double rt_prev, double z_prev, double e_out,
                             double rt_out, FlChainSmem& sm) {
    const int lane = threadIdx.x & 31;
    bool changed = false;
    uint32_t endmask;
    {
        const FlEWin w0 = fl_ewin_load(e, q, lane);
        const uint32_t nproc = fl_elev_nproc(w0, q, lane, endmask);
        fl_elev_window(e, w0, q, nproc, lane, root, rt_prev, z_prev, e_out, rt_out, changed, sm);
        if (endmask) return __ballot_sync(FL_FULL, changed) != 0u;
        q += 32u;
    }
    FlEWin ring[FL_EDEPTH];
#pragma unroll
    for (int j = 0; j < FL_EDEPTH; ++j) ring[j] = fl_ewin_load(e, q + 32u * (uint32_t)j, lane);
    for (;;) {
        const uint32_t nproc = fl_elev_nproc(ring[0], q, lane, endmask);
        const FlEWin cur = ring[0];
#pragma unroll
        for (int j = 0; j + 1 < FL_EDEPTH; ++j) ring[j] = ring[j + 1];
        if (!endmask) ring[FL_EDEPTH - 1] = fl_ewin_load(e, q + 32u * (uint32_t)FL_EDEPTH, lane);
        fl_elev_window(e, cur, q, nproc, lane, root, rt_prev, z_prev, e_out, rt_out, changed, sm);
        if (endmask) break;
        q += 32u;
    }
    return __ballot_sync(FL_FULL, changed) != 0u;
}
#endif

// One thread per segment head for the first 8 sites; segments that go on are finished by the whole warp,
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
            if (!ended) ended = fl_elev_batch<4>(e, q, h, false, root, rt_prev, z_prev, e_out, rt_out, changed);
            longseg = !ended;
#ifdef FL_EMU
            if (longseg) while (!fl_elev_batch<4>(e, q, h, false, root, rt_prev, z_prev, e_out, rt_out, changed)) {}
#endif
        }
    }
#ifndef FL_EMU
    __shared__ FlChainSmem chain_smem[4];  // one per warp (128 threads)
    FlChainSmem& sm = chain_smem[threadIdx.x >> 5];
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
        const bool ch = fl_elev_warp(e, q_s, root_s, rtp_s, zp_s, eo_s, ro_s, sm);
        if (lane == src) changed |= ch;
    }
#endif
    if (changed) e.flags[FL_FLAG_CHANGED] = 1u;
}

// Levels with few segments (the top of the forest: few, long segments): one WARP per segment, so that long
// segments of one level run side by side instead of one after the other inside a warp.
__global__ void __launch_bounds__(128) k_elev_flow_warps(uint32_t begin, uint32_t count,
                                                          const uint32_t* __restrict__ heads, FlElev e) {
#ifdef FL_EMU
    // emulation: one thread per segment, plain batches
    const uint32_t t = FL_TID;
    if (t >= count) return;
    const uint32_t h = heads[begin + t];
    const uint32_t p = e.recv[h];
    const bool is_root = (p == h);
    uint32_t root;
    double rt_prev, z_prev, e_out, rt_out;
    if (is_root) { root = e.is_outlet[h] ? h : FL_NONE; rt_prev = 0.0; z_prev = e.elev[h]; e_out = e.elev[h]; rt_out = 0.0; }
    else {
        root = e.root_of[p]; rt_prev = e.rt[p]; z_prev = e.elev[p];
        e_out = root != FL_NONE ? e.elev[root] : 0.0; rt_out = root != FL_NONE ? e.rt[root] : 0.0;
    }
    if (root == FL_NONE) {
        for (uint32_t r = h;; ++r) { e.root_of[r] = FL_NONE; if (r + 1u >= e.n || e.recv[r + 1u] != r) break; }
        return;
    }
    bool changed = false;
    uint32_t q = h;
    bool ended = fl_elev_batch<4>(e, q, h, is_root, root, rt_prev, z_prev, e_out, rt_out, changed);
    while (!ended) ended = fl_elev_batch<4>(e, q, h, false, root, rt_prev, z_prev, e_out, rt_out, changed);
    if (changed) e.flags[FL_FLAG_CHANGED] = 1u;
#else
    __shared__ FlChainSmem chain_smem[4];
    FlChainSmem& sm = chain_smem[threadIdx.x >> 5];
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= count) return;
    const uint32_t h = heads[begin + w];
    const uint32_t p = e.recv[h];
    const bool is_root = (p == h);
    uint32_t root;
    double rt_prev, z_prev, e_out, rt_out;
    if (is_root) { root = e.is_outlet[h] ? h : FL_NONE; rt_prev = 0.0; z_prev = e.elev[h]; e_out = e.elev[h]; rt_out = 0.0; }
    else {
        root = e.root_of[p]; rt_prev = e.rt[p]; z_prev = e.elev[p];
        e_out = root != FL_NONE ? e.elev[root] : 0.0; rt_out = root != FL_NONE ? e.rt[root] : 0.0;
    }
    if (root == FL_NONE) {
        if (lane == 0)
            for (uint32_t r = h;; ++r) { e.root_of[r] = FL_NONE; if (r + 1u >= e.n || e.recv[r + 1u] != r) break; }
        return;
    }
    // the head site, identically in every lane (the root's special cases live here); lane 0 stores
    bool changed = false;
    {
        const double t = e.tcel[h];
        const double eold = e.elev[h];
        const double rti = 0.0 + (rt_prev + t);
        if (is_root) rt_out = rti;
        double z = e_out + e.uplift[h] * fmax(rti - rt_out, 0.0);
        if (e.tan_slope) {
            const double ms = e.tan_slope[h];
            if (ms == ms) {
                const double d = e.drecv[h];
                const double slope = (z - z_prev) / d;
                if (slope > ms) z = z_prev + ms * d;
            }
        }
        changed = (z != eold);
        if (is_root) e_out = z;
        __syncwarp();  // every lane has read elev[h] before lane 0 overwrites it
        if (lane == 0) { e.elev[h] = z; e.rt[h] = rti; e.root_of[h] = root; e.lvl[h] = e.lvl_value; }
        rt_prev = rti;
        z_prev = z;
    }
    if (h + 1u < e.n && e.recv[h + 1u] == h)
        changed |= fl_elev_warp(e, h + 1u, root, rt_prev, z_prev, e_out, rt_out, sm);
    if (changed && lane == 0) e.flags[FL_FLAG_CHANGED] = 1u;
#endif
}

// The sparse levels at the top of the forest (few, long segments) in ONE launch: warps take the segments in level
// order through a ticket counter; a segment whose receiver lies on another segment waits until that segment's ticket is
// marked done.  The receiver's segment is on an earlier level, so its ticket was handed out earlier to a warp that is
// running: the wait always ends.  Saves a launch per level and the idle tail of every level.
struct FlFused {
    uint32_t count;            // segments in the fused prefix of the level order
    const uint32_t* heads;     // level order
    const uint32_t* seg_head;  // site -> head of its segment
    const uint32_t* ticket_of; // head -> position in the level order (valid for the fused prefix)
    uint32_t* done;            // per ticket
    uint32_t* next_ticket;
    const uint32_t* lvl_of;    // per ticket: the nesting height (value written to e.lvl)
};

__global__ void __launch_bounds__(256) k_fused_index(uint32_t count, const uint32_t* __restrict__ heads,
                                                      const uint32_t* __restrict__ hgt, uint32_t* __restrict__ ticket_of,
                                                      uint32_t* __restrict__ lvl_of, uint32_t* __restrict__ done) {
    const uint32_t i = FL_TID;
    if (i >= count) return;
    const uint32_t h = heads[i];
    ticket_of[h] = i;
    lvl_of[i] = hgt[h];
    done[i] = 0u;
}

__global__ void __launch_bounds__(128) k_elev_flow_fused(FlFused u, FlElev e) {
#ifdef FL_EMU
    // emulation: threads run one after the other in ticket order, so receivers are always finished
    const uint32_t t = FL_TID;
    if (t >= u.count) return;
    const uint32_t h = u.heads[t];
    const uint32_t p = e.recv[h];
    const bool is_root = (p == h);
    e.lvl_value = u.lvl_of[t];
    uint32_t root;
    double rt_prev, z_prev, e_out, rt_out;
    if (is_root) { root = e.is_outlet[h] ? h : FL_NONE; rt_prev = 0.0; z_prev = e.elev[h]; e_out = e.elev[h]; rt_out = 0.0; }
    else {
        root = e.root_of[p]; rt_prev = e.rt[p]; z_prev = e.elev[p];
        e_out = root != FL_NONE ? e.elev[root] : 0.0; rt_out = root != FL_NONE ? e.rt[root] : 0.0;
    }
    if (root == FL_NONE) {
        for (uint32_t r = h;; ++r) { e.root_of[r] = FL_NONE; if (r + 1u >= e.n || e.recv[r + 1u] != r) break; }
        return;
    }
    bool changed = false;
    uint32_t q = h;
    bool ended = fl_elev_batch<4>(e, q, h, is_root, root, rt_prev, z_prev, e_out, rt_out, changed);
    while (!ended) ended = fl_elev_batch<4>(e, q, h, false, root, rt_prev, z_prev, e_out, rt_out, changed);
    if (changed) e.flags[FL_FLAG_CHANGED] = 1u;
#else
    __shared__ FlChainSmem chain_smem[4];
    FlChainSmem& sm = chain_smem[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    for (;;) {
        uint32_t i = 0u;
        if (lane == 0) i = atomicAdd(u.next_ticket, 1u);
        i = __shfl_sync(FL_FULL, i, 0);
        if (i >= u.count) return;
        const uint32_t h = u.heads[i];
        const uint32_t p = e.recv[h];
        const bool is_root = (p == h);
        e.lvl_value = u.lvl_of[i];
        if (!is_root) {  // wait for the receiver's segment
            const uint32_t j = u.ticket_of[u.seg_head[p]];
            // every lane runs its own acquire load of the flag (one broadcast transaction per round)
            uint32_t spins = 0u;
            while (!__all_sync(FL_FULL, fl_ld_acquire(&u.done[j]) != 0u)) {
                __nanosleep(40);
                if (++spins > (1u << 24)) { if (lane == 0) atomicOr(&e.flags[FL_FLAG_BROKEN], 2u); break; }
            }
        }
        uint32_t root;
        double rt_prev, z_prev, e_out, rt_out;
        if (is_root) { root = e.is_outlet[h] ? h : FL_NONE; rt_prev = 0.0; z_prev = e.elev[h]; e_out = e.elev[h]; rt_out = 0.0; }
        else {  // written by another warp of this launch: read through L2
            root = fl_ld_cg(&e.root_of[p]); rt_prev = fl_ld_cg(&e.rt[p]); z_prev = fl_ld_cg(&e.elev[p]);
            e_out = root != FL_NONE ? fl_ld_cg(&e.elev[root]) : 0.0; rt_out = root != FL_NONE ? fl_ld_cg(&e.rt[root]) : 0.0;
        }
        bool changed = false;
        if (root == FL_NONE) {
            if (lane == 0)
                for (uint32_t r = h;; ++r) { e.root_of[r] = FL_NONE; if (r + 1u >= e.n || e.recv[r + 1u] != r) break; }
        } else {
            // the head site, identically in every lane (the root's special cases live here); lane 0 stores
            const double t = e.tcel[h];
            const double eold = e.elev[h];
            const double rti = 0.0 + (rt_prev + t);
            if (is_root) rt_out = rti;
            double z = e_out + e.uplift[h] * fmax(rti - rt_out, 0.0);
            if (e.tan_slope) {
                const double ms = e.tan_slope[h];
                if (ms == ms) {
                    const double d = e.drecv[h];
                    const double slope = (z - z_prev) / d;
                    if (slope > ms) z = z_prev + ms * d;
                }
            }
            changed = (z != eold);
            if (is_root) e_out = z;
            __syncwarp();  // every lane has read elev[h] before lane 0 overwrites it
            if (lane == 0) { e.elev[h] = z; e.rt[h] = rti; e.root_of[h] = root; e.lvl[h] = e.lvl_value; }
            rt_prev = rti;
            z_prev = z;
            if (h + 1u < e.n && e.recv[h + 1u] == h)
                changed |= fl_elev_warp(e, h + 1u, root, rt_prev, z_prev, e_out, rt_out, sm);
            if (changed && lane == 0) e.flags[FL_FLAG_CHANGED] = 1u;
        }
        __syncwarp();  // the lanes' stores happen before lane 0's release fence
        if (lane == 0) { fl_fence_release(); atomicExch(&u.done[i], 1u); }
    }
#endif
}

// keys for sorting segment heads by descending nesting height: key = base - hgt (heads), FL_NONE otherwise
__global__ void __launch_bounds__(256) k_flow_sort_keys(uint32_t n, const uint32_t* __restrict__ hgt, uint32_t base,
                                                         uint32_t* __restrict__ keys, uint32_t* flags) {
    uint32_t q = FL_TID;
    if (q >= n) return;
    if (q == 0u) flags[FL_FLAG_K4MAXH] = flags[FL_FLAG_MAXDEPTH];  // K4's result, before the slot is reused below
    const uint32_t h = hgt[q];
    uint32_t key = FL_NONE;
    if (h != FL_NONE) {
        if (h > base) atomicOr(&flags[FL_FLAG_BROKEN], 4u);
        else key = base - h;
    }
    keys[q] = key;
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
