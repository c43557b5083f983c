// fl_flow.cuh -- "sweep" = 3: dataflow tree sweeps on DYNAMIC segments (the default).
//
// A segment is a maximal run of positions q, q+1, ... with recv[q+1] == q: a piece of a flow path that is
// contiguous in the current site numbering.  Right after a layout rebuild (fl_paths.cuh, heavy paths) the
// segments are the heavy paths; when receivers change in later iterations a chain simply breaks into
// shorter segments -- nothing has to be patched, so a numbering is kept for many iterations and rebuilt
// only when it has degraded.  Every floating-point addition below is one the reference performs, in the
// reference's order (generator.rs:154-174); only WHO performs it and WHEN is reorganised.
//
// K4 drainage area, bottom-up, two launches and no level structure:
//   pass 1 (k_area_flow): every leaf starts a THREAD-level flow that climbs its segment with the running
//     area in a register.  A flow that finishes a segment head h reports to the parent site p = recv[h]; the
//     LAST child to report ("last arriver") gathers p's non-chain children in reverse adjacency order into
//        pre = a_p + (children before the chain child),   post1, post2 = (children after it)
//     and either resumes the flow that is waiting below p or leaves the values for the flow still to come.
//     Nobody spins: a thread continues with work that is ready or exits.  A flow that has climbed
//     `park_after` sites in a row is on a long chain: it parks (stores its running value) and exits.
//   pass 2 (k_area_flow_long): the parked flows are continued by whole WARPS: 32-site windows of the chain
//     are fetched by the lanes together (coalesced, prefetched one window ahead) and only the serial
//     additions run lane to lane.  A warp stays a warp through all later hand-offs of its flow.
//   The pass also yields, per segment head, the nesting height (longest chain of hand-offs below it), which
//   orders the top-down sweep exactly for the CURRENT forest.
// K5 response time / elevation, top-down: one launch per nesting height (k_elev_flow), one thread per
//   segment for its first sites, the rest of long segments by the whole warp.
#pragma once
#include "fl_paths.cuh"

// state word of a site: [0,8) children reported, [8,12) np (posts; 15 = more than two), [12,30) hpre (max child
// height + 1; 0x3FFFF = too large, read hpre[]), bit 30 pre/posts published, bit 31 a flow is waiting below
#define FL_ST_COUNT_MASK 0x000000FFu
#define FL_ST_NP_SHIFT 8
#define FL_ST_NP_MASK 0x00000F00u
#define FL_ST_HP_SHIFT 12
#define FL_ST_HP_MASK 0x3FFFF000u
#define FL_ST_HP_OVER 0x3FFFFu
#define FL_ST_PRE_READY 0x40000000u
#define FL_ST_SCAN_ARRIVED 0x80000000u
#define FL_TB 4  // sites per batch of a thread-level flow

#ifdef FL_EMU
template <class T> __device__ __forceinline__ T fl_ld_cg(const T* p) { return *p; }
__device__ __forceinline__ uint32_t fl_ld_relaxed(const uint32_t* p) { return *p; }
__device__ __forceinline__ uint32_t fl_dep0(uint32_t) { return 0u; }
#else
template <class T> __device__ __forceinline__ T fl_ld_cg(const T* p) { return __ldcg(p); }
// relaxed gpu-scope load of a flag word (served by L2, pipelinable)
__device__ __forceinline__ uint32_t fl_ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// An opaque zero derived from v.  Adding it to an index makes the following load ADDRESS-DEPENDENT on the
// flag (or atomic result) v, so it is issued only after v has arrived and is served by L2 after the
// writer's fence + flag update: the reader side needs no fence of its own.
__device__ __forceinline__ uint32_t fl_dep0(uint32_t v) {
    uint32_t z;
    asm volatile("and.b32 %0, %1, 0;" : "=r"(z) : "r"(v));
    return z;
}
#endif

__device__ __forceinline__ uint32_t fl_st_publish(uint32_t np, uint32_t hp) {
    const uint32_t h18 = hp < FL_ST_HP_OVER ? hp : FL_ST_HP_OVER;
    return FL_ST_PRE_READY | (np << FL_ST_NP_SHIFT) | (h18 << FL_ST_HP_SHIFT);
}
__device__ __forceinline__ uint32_t fl_st_np(uint32_t s) { return (s & FL_ST_NP_MASK) >> FL_ST_NP_SHIFT; }
// max child height + 1 of site i, from its state word s (i must already carry the address dependency on s)
__device__ __forceinline__ uint32_t fl_st_hp(uint32_t s, const uint32_t* hpre, uint32_t i) {
    const uint32_t h = (s & FL_ST_HP_MASK) >> FL_ST_HP_SHIFT;
    return h != FL_ST_HP_OVER ? h : fl_ld_cg(&hpre[i]);
}
// pre[] and post1[] are filled with this bit pattern (a NaN no addition produces) before every K4, so a
// speculative load of them validates itself: anything else is a value published in this iteration.
__device__ __forceinline__ bool fl_is_unset(double v) {
#ifdef FL_EMU
    unsigned long long b; std::memcpy(&b, &v, 8); return b == 0xFFFFFFFFFFFFFFFFull;
#else
    return __double_as_longlong(v) == -1ll;
#endif
}

struct FlFlow {
    uint32_t n;
    const uint32_t* row_ptr;
    const uint32_t* col;
    const uint32_t* recv;
    const uint32_t* cmask;
    const double* areas;
    double* A;
    const uint32_t* nwait;  // non-leaf non-chain children per site (k_count_waits); leaves never report
    uint32_t* state;  // zeroed before the launch: [0,24) children reported, [24,28) np, bit 30 pre ready, bit 31 flow waiting
    double* pre;
    double* post1;
    double* post2;
    double* xbuf;     // running area handed over at a waiting / parked site
    uint32_t* hbuf;   // running nesting height handed over with it
    uint32_t* hgt;    // out: nesting height for segment heads, FL_NONE elsewhere
    uint32_t* hpre;   // max (height + 1) over the non-chain children of a site, written with pre
    uint32_t* flags;
    uint32_t* parked;    // sites where a long flow was parked for the warp-level pass
    uint32_t* counters;  // [0] = number of parked flows, [1] = next one to take
    uint32_t park_after; // a thread parks its flow after climbing this many sites in a row (0 = never)
    uint32_t* next_list;   // round-synchronous mode: sites whose last child has just reported
    uint32_t* next_count;
    const uint32_t* lvl;   // nesting height of each site's segment in the PREVIOUS iteration (schedule hint)
    uint32_t level, top_level;  // k_area_flow_long: take parked flows with min(lvl, top_level) == level
};

// per-warp staging area of the serial chains (warp-level scans)
struct FlChainSmem { double in[32]; double out[32]; double aux1[32]; double aux2[32]; };

// non-chain children of p, reverse adjacency order -> pre / posts; returns np (15 = more than two posts).
// Leaf children never run a flow: their area is their own cell area and their height 0.  `dep` is an opaque
// zero that orders the loads of the other children's results after the atomic that made us last arriver.
__device__ __forceinline__ uint32_t fl_gather_lights(const FlFlow& f, uint32_t p, bool has_chain, uint32_t dep,
                                                     double& pre, double& p1, double& p2, uint32_t& hmax) {
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
        double v;
        uint32_t hc = 1u;
        if (f.cmask[c] == 0u) {
            v = f.areas[c];
        } else {
            v = fl_ld_cg(&f.A[c + dep]);
            hc = fl_ld_cg(&f.hgt[c + dep]) + 1u;
        }
        if (hc > hmax) hmax = hc;
        if (!seen) pre += v;
        else { if (np == 0) p1 = v; else if (np == 1) p2 = v; ++np; }
    }
    return np > 2 ? 15u : np;
}

// slow path for np == 15: add every child after the chain child, in order
__device__ double fl_add_posts(const FlFlow& f, uint32_t p, double y) {
    bool seen = false;
    const uint32_t s0 = f.row_ptr[p];
    uint32_t m = f.cmask[p];
    while (m) {
        const uint32_t b = 31u - (uint32_t)__clz((int)m);
        m ^= 1u << b;
        const uint32_t c = f.col[s0 + b];
        if (c == p + 1u) { seen = true; continue; }
        if (seen) y += (f.cmask[c] == 0u) ? f.areas[c] : fl_ld_cg(&f.A[c]);
    }
    return y;
}

// ------------------------------------------------------------------------------------------------
// thread-level flow
// ------------------------------------------------------------------------------------------------
// Entry states: at_head = false: about to process site `cur` (x = area of its chain child if has_chain);
//               at_head = true : site `cur` is a finished segment head with area y and receiver p.
//               resume  = true : pre/posts of `cur` are passed in registers (the caller is its last arriver).
// defer = false: a flow that becomes last arriver at a site continues there at once (pure dataflow).
// defer = true : it returns that site instead (round-synchronous mode: the caller queues it for the next launch,
//                and the kernel boundary publishes the children's results, so no fence is needed to report).
struct FlPre { double pre, p1, p2; uint32_t np, hp; };

__device__ uint32_t fl_flow_thread(const FlFlow& f, uint32_t cur, double x, uint32_t hrun, bool has_chain, bool at_head,
                                   double y, uint32_t p, bool may_park, bool defer, bool resume, FlPre in) {
    uint32_t climbed = 0;  // sites climbed through the fast path without a break
    double pre = in.pre, p1 = in.p1, p2 = in.p2;
    uint32_t np = in.np, hp = in.hp;
    for (;;) {
        if (!at_head && !resume) {
            // ---- fast path: up to FL_TB consecutive sites, loads issued up front ----
            const uint32_t nb = cur + 1u < (uint32_t)FL_TB ? cur + 1u : (uint32_t)FL_TB;
            uint32_t cm[FL_TB], rc[FL_TB], st[FL_TB], hq[FL_TB];
            double ar[FL_TB], pr[FL_TB], q1[FL_TB], q2[FL_TB];
#pragma unroll
            for (int k = 0; k < FL_TB; ++k) {
                cm[k] = 0u; rc[k] = FL_NONE; ar[k] = 0.0;
                if ((uint32_t)k < nb) { cm[k] = f.cmask[cur - k]; rc[k] = f.recv[cur - k]; ar[k] = f.areas[cur - k]; }
            }
#pragma unroll
            for (int k = 0; k < FL_TB; ++k) {
                st[k] = 0u;
                const uint32_t nl = (uint32_t)__popc(cm[k]) - ((k > 0 || has_chain) ? 1u : 0u);
                if ((uint32_t)k < nb && cm[k] != 0u && nl > 0u) st[k] = fl_ld_relaxed(&f.state[cur - k]);
            }
#pragma unroll
            for (int k = 0; k < FL_TB; ++k) {
                pr[k] = 0.0; q1[k] = 0.0; q2[k] = 0.0; hq[k] = 0u;
                if (st[k] & FL_ST_PRE_READY) {
                    const uint32_t npk = fl_st_np(st[k]);
                    const uint32_t i = cur - k + fl_dep0(st[k]);
                    pr[k] = fl_ld_cg(&f.pre[i]);
                    hq[k] = fl_st_hp(st[k], f.hpre, i);
                    if (npk >= 1u && npk != 15u) q1[k] = fl_ld_cg(&f.post1[i]);
                    if (npk >= 2u && npk != 15u) q2[k] = fl_ld_cg(&f.post2[i]);
                }
            }
            uint32_t done = 0;  // sites of the batch finished and climbed past
#pragma unroll
            for (int k = 0; k < FL_TB; ++k) {
                if (!at_head && done == (uint32_t)k && (uint32_t)k < nb) {
                    const uint32_t idx = cur - k;
                    const bool hc = (k > 0) || has_chain;
                    const uint32_t nl = (uint32_t)__popc(cm[k]) - (hc ? 1u : 0u);
                    bool ok = true;
                    if (nl == 0u) {
                        y = hc ? (ar[k] + x) : ar[k];
                    } else if (st[k] & FL_ST_PRE_READY) {
                        const uint32_t npk = fl_st_np(st[k]);
                        y = hc ? (pr[k] + x) : pr[k];
                        if (npk == 15u) y = fl_add_posts(f, idx, y);
                        else {
                            if (npk >= 1u) y += q1[k];
                            if (npk >= 2u) y += q2[k];
                        }
                        if (hq[k] > hrun) hrun = hq[k];
                    } else {
                        ok = false;  // children of idx still running: hand-off path below
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
                    f.xbuf[cur] = x;  // a long chain: leave it to the warp-level pass
                    f.hbuf[cur] = hrun;
                    f.parked[atomicAdd(&f.counters[0], 1u)] = cur;
                    return FL_NONE;
                }
                continue;
            }
            climbed = 0;
        }
        if (!at_head) {
            // ---- one site: wait for / take over from its children ----
            const uint32_t nlight = (uint32_t)__popc(f.cmask[cur]) - (has_chain ? 1u : 0u);
            if (nlight == 0u) {
                y = has_chain ? (f.areas[cur] + x) : f.areas[cur];
            } else {
                if (!resume) {
                    uint32_t s = fl_ld_relaxed(&f.state[cur]);
                    if (!(s & FL_ST_PRE_READY)) {
                        f.xbuf[cur] = x;
                        f.hbuf[cur] = hrun;
                        __threadfence();
                        s = atomicOr(&f.state[cur], FL_ST_SCAN_ARRIVED);
                        if (!(s & FL_ST_PRE_READY)) return FL_NONE;  // the last arriver of `cur` takes over
                    }
                    const uint32_t i = cur + fl_dep0(s);
                    np = fl_st_np(s);
                    pre = fl_ld_cg(&f.pre[i]);
                    hp = fl_st_hp(s, f.hpre, i);
                    if (np >= 1u && np != 15u) p1 = fl_ld_cg(&f.post1[i]);
                    if (np >= 2u && np != 15u) p2 = fl_ld_cg(&f.post2[i]);
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
        at_head = false;
        f.hgt[cur] = hrun;
        if (p == cur) {  // tree root: its segment has the largest nesting height of the tree
            if (hrun > 0u) atomicMax(&f.flags[FL_FLAG_MAXDEPTH], hrun);
            return FL_NONE;
        }
        if (!defer) __threadfence();  // publish A[cur], hgt[cur] (deferred mode: the kernel boundary does)
        const uint32_t prev = atomicAdd(&f.state[p], 1u);
        const uint32_t arrived = (prev & FL_ST_COUNT_MASK) + 1u;
        if (arrived < f.nwait[p]) return FL_NONE;
        if (defer) return p;  // queue p: its last arriver runs in the next launch
        // last arriver at p: gather p's non-chain children
        const bool p_has_chain = (p + 1u < f.n) && (f.recv[p + 1u] == p);
        np = fl_gather_lights(f, p, p_has_chain, fl_dep0(prev), pre, p1, p2, hp);
        if (!p_has_chain) {  // p ends its segment: nobody climbs into it, the flow continues here
            cur = p; has_chain = false; x = 0.0; hrun = 0; resume = true; climbed = 0;
            continue;
        }
        f.pre[p] = pre;
        f.hpre[p] = hp;
        if (np >= 1u && np != 15u) f.post1[p] = p1;
        if (np >= 2u && np != 15u) f.post2[p] = p2;
        __threadfence();
        const uint32_t old = atomicOr(&f.state[p], fl_st_publish(np, hp));
        if (!(old & FL_ST_SCAN_ARRIVED)) return FL_NONE;  // the flow below p has not arrived yet; it will pick these up
        const uint32_t pi = p + fl_dep0(old);
        x = fl_ld_cg(&f.xbuf[pi]);
        hrun = fl_ld_cg(&f.hbuf[pi]);
        cur = p; has_chain = true; resume = true; climbed = 0;
    }
}

// pre-pass A: how many children will REPORT to each site = its non-leaf segment heads.
// (Leaves are resolved by their parent directly; chain children hand over inside the segment.)
__global__ void __launch_bounds__(256) k_count_waits(uint32_t n, const uint32_t* __restrict__ recv,
                                                      const uint32_t* __restrict__ cmask, uint32_t* nwait) {
    const uint32_t q = FL_TID;
    if (q >= n) return;
    if (cmask[q] == 0u) return;
    const uint32_t p = recv[q];
    if (p == q || (q > 0u && p == q - 1u)) return;  // root, or chained to its receiver
    atomicAdd(&nwait[p], 1u);
}

// pre-pass B: sites whose non-chain children are all leaves need nobody: their pre / posts are gathered
// here, in parallel, and published as ready (plain stores: the kernel boundary orders them).
__global__ void __launch_bounds__(256) k_simple_pre(FlFlow f) {
    const uint32_t q = FL_TID;
    if (q >= f.n) return;
    const uint32_t cm = f.cmask[q];
    if (cm == 0u || f.nwait[q] != 0u) return;
    const bool has_chain = (q + 1u < f.n) && (f.recv[q + 1u] == q);
    if ((uint32_t)__popc(cm) - (has_chain ? 1u : 0u) == 0u) return;  // only the chain child
    double pre, p1, p2;
    uint32_t hp;
    const uint32_t np = fl_gather_lights(f, q, has_chain, 0u, pre, p1, p2, hp);
    f.pre[q] = pre;
    f.hpre[q] = hp;
    if (np >= 1u && np != 15u) f.post1[q] = p1;
    if (np >= 2u && np != 15u) f.post2[q] = p2;
    f.state[q] = fl_st_publish(np, hp);
}

// pass 1: thread-level flows start at (a) leaves that are the tail of a chain and (b) ready sites that end
// their segment.  All other leaves only publish their own area and height.
__global__ void __launch_bounds__(256) k_area_flow(FlFlow f) {
    const uint32_t q = FL_TID;
    if (q >= f.n) return;
    const uint32_t cm = f.cmask[q];
    if (cm == 0u) {
        const double y = f.areas[q];
        const uint32_t p = f.recv[q];
        f.A[q] = y;
        if (q > 0u && p == q - 1u) {  // the leaf is the tail of a chain: climb
            f.hgt[q] = FL_NONE;
            fl_flow_thread(f, q - 1u, y, 0u, true, false, 0.0, FL_NONE, true, false, false, FlPre{0.0, 0.0, 0.0, 0u, 0u});
        } else {
            f.hgt[q] = 0u;  // a one-site segment; its parent reads areas[q] itself
        }
        return;
    }
    if (f.nwait[q] != 0u) return;                             // children will report: the last one continues here
    if ((q + 1u < f.n) && (f.recv[q + 1u] == q)) return;      // a chain child will climb into q
    fl_flow_thread(f, q, 0.0, 0u, false, false, 0.0, FL_NONE, true, false, false, FlPre{0.0, 0.0, 0.0, 0u, 0u});  // tail whose children are all leaves
}

// ------------------------------------------------------------------------------------------------
// K4, round-synchronous variant (option "k4_rounds", default): the same protocol, but a flow that becomes the
// last arriver at a site does not continue there -- the site is queued and processed by the next launch.
// Every launch is then a regular kernel over a compact list (no long-lived divergent threads); the number of
// launches is the nesting height of the segment forest.  Flows on long chains still park; the parked flows are
// finished by k_area_flow_long (warp-level dataflow) after the last round.
// ------------------------------------------------------------------------------------------------
// append `value` (if pred) to list[*count ...]: one atomic per warp
__device__ __forceinline__ void fl_append(uint32_t* list, uint32_t* count, uint32_t value, bool pred) {
#ifdef FL_EMU
    if (pred) list[atomicAdd(count, 1u)] = value;
#else
    const uint32_t mask = __ballot_sync(FL_FULL, pred);
    if (!mask) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs((int)mask) - 1;
    uint32_t base = 0u;
    if (lane == leader) base = atomicAdd(count, (uint32_t)__popc(mask));
    base = __shfl_sync(FL_FULL, base, leader);
    if (pred) list[base + (uint32_t)__popc(mask & ((1u << lane) - 1u))] = value;
#endif
}

// round 0 list + everything that needs no flow: leaves publish their own area / height, sites whose
// non-chain children are all leaves get their pre / posts (as k_simple_pre), segment tails that wait for
// nobody are queued as starters.
__global__ void __launch_bounds__(256) k_flow_prepare(FlFlow f, uint32_t* list, uint32_t* count) {
    const uint32_t q = FL_TID;
    bool start = false;
    if (q < f.n) {
        const uint32_t cm = f.cmask[q];
        const bool has_chain = (q + 1u < f.n) && (f.recv[q + 1u] == q);
        if (cm == 0u) {
            const uint32_t p = f.recv[q];
            if (q > 0u && p == q - 1u) {
                start = true;  // leaf at the tail of a chain: climbs in round 0
            } else {
                f.A[q] = f.areas[q];  // one-site segment; its parent reads areas[q] itself
                f.hgt[q] = 0u;
            }
        } else if (f.nwait[q] == 0u) {
            if ((uint32_t)__popc(cm) - (has_chain ? 1u : 0u) != 0u) {
                double pre, p1, p2;
                uint32_t hp;
                const uint32_t np = fl_gather_lights(f, q, has_chain, 0u, pre, p1, p2, hp);
                f.pre[q] = pre;
                f.hpre[q] = hp;
                if (np >= 1u && np != 15u) f.post1[q] = p1;
                if (np >= 2u && np != 15u) f.post2[q] = p2;
                f.state[q] = fl_st_publish(np, hp);
            }
            start = !has_chain;  // a tail that waits for nobody
        }
    }
    fl_append(list, count, q, start);
}

// one queued site: either a starter (tail that waits for nobody) or a site whose last child has reported
__device__ uint32_t fl_round_entry(const FlFlow& f, uint32_t p) {
    const bool has_chain = (p + 1u < f.n) && (f.recv[p + 1u] == p);
    const uint32_t nl = (uint32_t)__popc(f.cmask[p]) - (has_chain ? 1u : 0u);
    FlPre in{0.0, 0.0, 0.0, 0u, 0u};
    if (nl == 0u || (f.state[p] & FL_ST_PRE_READY))  // leaf tail, or prepared by k_flow_prepare
        return fl_flow_thread(f, p, 0.0, 0u, false, false, 0.0, FL_NONE, true, true, false, in);
    // we are p's last arriver (its children reported in earlier launches)
    in.np = fl_gather_lights(f, p, has_chain, 0u, in.pre, in.p1, in.p2, in.hp);
    if (!has_chain) return fl_flow_thread(f, p, 0.0, 0u, false, false, 0.0, FL_NONE, true, true, true, in);
    f.pre[p] = in.pre;
    f.hpre[p] = in.hp;
    if (in.np >= 1u && in.np != 15u) f.post1[p] = in.p1;
    if (in.np >= 2u && in.np != 15u) f.post2[p] = in.p2;
    __threadfence();
    const uint32_t old = atomicOr(&f.state[p], fl_st_publish(in.np, in.hp));
    if (!(old & FL_ST_SCAN_ARRIVED)) return FL_NONE;  // the flow below p has not arrived yet; it will pick these up
    const uint32_t pi = p + fl_dep0(old);
    const double x = fl_ld_cg(&f.xbuf[pi]);
    const uint32_t hrun = fl_ld_cg(&f.hbuf[pi]);
    return fl_flow_thread(f, p, x, hrun, true, false, 0.0, FL_NONE, true, true, true, in);
}

__global__ void __launch_bounds__(256) k_area_round(FlFlow f, const uint32_t* __restrict__ list,
                                                     const uint32_t* count_in, uint32_t* next, uint32_t* count_out) {
    const uint32_t cnt = fl_ld_cg(count_in);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x; base < cnt; base += stride) {  // block-uniform trip count
        const uint32_t i = base + threadIdx.x;
        uint32_t out = FL_NONE;
        if (i < cnt) out = fl_round_entry(f, list[i]);
        fl_append(next, count_out, out, out != FL_NONE);
    }
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
    double ar, pre, p1;  // pre / p1: speculative copies of pre[] / post1[] (self-validating, see fl_is_unset)
};

__device__ __forceinline__ FlWin fl_win_load(const FlFlow& f, uint32_t base, int lane) {
    FlWin w;
    w.cm = 0u; w.rc = FL_NONE; w.st = 0u; w.ar = 0.0; w.pre = 0.0; w.p1 = 0.0;
    const long long li = (long long)base - lane;
    if (li >= 0) {
        const uint32_t idx = (uint32_t)li;
        w.cm = f.cmask[idx];
        w.rc = f.recv[idx];
        w.ar = f.areas[idx];
        w.st = fl_ld_relaxed(&f.state[idx]);
        w.pre = fl_ld_cg(&f.pre[idx]);
        w.p1 = fl_ld_cg(&f.post1[idx]);
    }
    return w;
}

#define FL_WDEPTH 3  // windows kept in flight on a long chain

__device__ void fl_flow_warp(const FlFlow& f, uint32_t cur, double x, uint32_t hrun, bool has_chain, FlChainSmem& sm,
                             bool defer) {
    const int lane = threadIdx.x & 31;
    bool resume = false;
    double pre = 0.0, p1 = 0.0, p2 = 0.0;
    uint32_t np = 0, hp = 0;
    // ring[j] holds the window whose lane 0 is site cur - 32*j, for j < nring
    FlWin ring[FL_WDEPTH];
    uint32_t nring = 0;
#pragma unroll
    for (int j = 0; j < FL_WDEPTH; ++j) { ring[j].cm = 0u; ring[j].rc = FL_NONE; ring[j].st = 0u; ring[j].ar = 0.0; ring[j].pre = 0.0; ring[j].p1 = 0.0; }
    for (;;) {
        double y = 0.0;
        uint32_t p = FL_NONE;
        bool at_head = false;
        if (!resume) {
            if (nring == 0u) { ring[0] = fl_win_load(f, cur, lane); nring = 1u; }
            const FlWin win = ring[0];
            const long long li = (long long)cur - lane;
            const bool valid = li >= 0;
            const uint32_t idx = (uint32_t)li;
            const uint32_t rc_up = __shfl_up_sync(FL_FULL, win.rc, 1);
            const bool link = valid && (lane == 0 || rc_up == idx);
            const uint32_t linkmask = __ballot_sync(FL_FULL, link);
            const uint32_t nchain = linkmask == FL_FULL ? 32u : (uint32_t)__ffs((int)~linkmask) - 1u;
            const bool hc = (lane > 0) || has_chain;
            const bool inwin = (uint32_t)lane < nchain;
            const uint32_t nl = inwin ? (uint32_t)__popc(win.cm) - (hc ? 1u : 0u) : 0u;
            const bool ready = inwin && (nl == 0u || (win.st & FL_ST_PRE_READY));
            const uint32_t readymask = __ballot_sync(FL_FULL, ready);
            const uint32_t nproc = readymask == FL_FULL ? 32u : (uint32_t)__ffs((int)~readymask) - 1u;
            const bool lit = (uint32_t)lane < nproc && nl > 0u;
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
                q1 = win.p1;
                hq = (win.st & FL_ST_HP_MASK) >> FL_ST_HP_SHIFT;
            }
            // speculative copies that cannot be trusted (or are not enough) are re-read, address-dependent on the flag
            const bool redo = lit && (fl_is_unset(b) || (npk >= 1u && npk != 15u && fl_is_unset(q1)) ||
                                      (npk >= 2u && npk != 15u) || hq == FL_ST_HP_OVER);
            if (__ballot_sync(FL_FULL, redo)) {
                if (redo) {
                    const uint32_t i = idx + fl_dep0(win.st);
                    b = fl_ld_cg(&f.pre[i]);
                    hq = fl_st_hp(win.st, f.hpre, i);
                    if (npk >= 1u && npk != 15u) q1 = fl_ld_cg(&f.post1[i]);
                    if (npk >= 2u && npk != 15u) q2 = fl_ld_cg(&f.post2[i]);
                }
            }
            const uint32_t postmask = __ballot_sync(FL_FULL, lit && npk != 0u);
            double mine;
            if (postmask == 0u) {
                // common case, no children after the chain child anywhere in the window: y_k = b_k + y_{k-1}.
                // Terms staged in shared memory, identical chain in every lane over broadcast reads.
                __syncwarp();
                sm.in[lane] = b;
                __syncwarp();
                double r = x;
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const double bk = sm.in[k];
                    if ((uint32_t)k < nproc) r = ((k > 0) || has_chain) ? (bk + r) : bk;
                    sm.out[k] = r;
                }
                x = r;
                __syncwarp();
                mine = sm.out[lane];
            } else {
                mine = 0.0;
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const double bk = fl_shfl(b, k);
                    if ((uint32_t)k < nproc) {
                        double yy = ((k > 0) || has_chain) ? (bk + x) : bk;
                        if ((postmask >> k) & 1u) {  // children after the chain child
                            const uint32_t npk_k = __shfl_sync(FL_FULL, npk, k);
                            if (npk_k == 15u) {
                                double v = 0.0;
                                if (lane == k) v = fl_add_posts(f, idx, yy);
                                __syncwarp();
                                yy = fl_shfl(v, k);
                            } else {
                                const double q1k = fl_shfl(q1, k);
                                const double q2k = fl_shfl(q2, k);
                                yy += q1k;
                                if (npk_k >= 2u) yy += q2k;
                            }
                        }
                        x = yy;
                        if (lane == k) mine = yy;
                    }
                }
            }
            const uint32_t hw = fl_warp_max((uint32_t)lane < nproc ? hq : 0u);
            if (hw > hrun) hrun = hw;
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
                    p = __shfl_sync(FL_FULL, win.rc, lastl);
                    cur -= (uint32_t)lastl;
                    nring = 0u;
                } else {
                    has_chain = true;
                    cur -= nproc;
                    if (nproc == 32u) {  // whole window climbed: shift the ring
#pragma unroll
                        for (int j = 0; j + 1 < FL_WDEPTH; ++j) ring[j] = ring[j + 1];
                        nring = nring > 0u ? nring - 1u : 0u;
                        continue;
                    }
                    nring = 0u;
                }
            } else {
                nring = 0u;
            }
        }
        if (!at_head) {
            // one site `cur` (uniform across the warp; lane 0 does the side effects)
            const uint32_t nlight = (uint32_t)__popc(f.cmask[cur]) - (has_chain ? 1u : 0u);
            if (nlight == 0u) {
                y = has_chain ? (f.areas[cur] + x) : f.areas[cur];
            } else {
                if (!resume) {
                    uint32_t sv = 0u;
                    if (lane == 0) {
                        sv = fl_ld_relaxed(&f.state[cur]);
                        if (!(sv & FL_ST_PRE_READY)) {
                            f.xbuf[cur] = x;
                            f.hbuf[cur] = hrun;
                            __threadfence();
                            sv = atomicOr(&f.state[cur], FL_ST_SCAN_ARRIVED);
                        }
                    }
                    sv = __shfl_sync(FL_FULL, sv, 0);
                    if (!(sv & FL_ST_PRE_READY)) return;  // the last arriver of `cur` takes over
                    const uint32_t i = cur + fl_dep0(sv);
                    np = fl_st_np(sv);
                    pre = fl_ld_cg(&f.pre[i]);
                    hp = fl_st_hp(sv, f.hpre, i);
                    if (np >= 1u && np != 15u) p1 = fl_ld_cg(&f.post1[i]);
                    if (np >= 2u && np != 15u) p2 = fl_ld_cg(&f.post2[i]);
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
            nring = 0u;
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
        uint32_t prev = 0u;
        if (lane == 0) {
            if (!defer) __threadfence();
            prev = atomicAdd(&f.state[p], 1u);
        }
        prev = __shfl_sync(FL_FULL, prev, 0);
        const uint32_t arrived = (prev & FL_ST_COUNT_MASK) + 1u;
        if (arrived < f.nwait[p]) return;
        if (defer) {  // queue p for the next launch
            if (lane == 0) f.next_list[atomicAdd(f.next_count, 1u)] = p;
            return;
        }
        const bool p_has_chain = (p + 1u < f.n) && (f.recv[p + 1u] == p);
        np = fl_gather_lights(f, p, p_has_chain, fl_dep0(prev), pre, p1, p2, hp);  // uniform: every lane, same values
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
            old = atomicOr(&f.state[p], fl_st_publish(np, hp));
        }
        old = __shfl_sync(FL_FULL, old, 0);
        if (!(old & FL_ST_SCAN_ARRIVED)) return;
        const uint32_t pi = p + fl_dep0(old);
        x = fl_ld_cg(&f.xbuf[pi]);
        hrun = fl_ld_cg(&f.hbuf[pi]);
        cur = p; has_chain = true; resume = true;
    }
}
#endif

// pass 2: the parked (long) flows, continued by whole warps.  With a schedule hint (f.lvl != null) the
// kernel is launched once per level of the PREVIOUS iteration's nesting height, bottom-up, and takes only the
// parked flows of that level: when the forest has not changed, every flow then finds all its tributaries
// finished and scans its chain without a single wait.  A wrong hint costs hand-offs, never correctness.
__global__ void __launch_bounds__(256, 2) k_area_flow_long(FlFlow f) {
    const uint32_t count = fl_ld_cg(&f.counters[0]);
#ifdef FL_EMU
    for (uint32_t i = FL_TID; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t cur = f.parked[i];
        if (f.lvl) {
            const uint32_t l = f.lvl[cur] < f.top_level ? f.lvl[cur] : f.top_level;
            if (l != f.level) continue;
        }
        const bool defer = f.next_list != nullptr;
        const uint32_t out = fl_flow_thread(f, cur, f.xbuf[cur], f.hbuf[cur], true, false, 0.0, FL_NONE, false, defer, false,
                                            FlPre{0.0, 0.0, 0.0, 0u, 0u});
        if (out != FL_NONE) f.next_list[atomicAdd(f.next_count, 1u)] = out;
    }
#else
    __shared__ FlChainSmem chain_smem[8];  // one per warp (256 threads)
    FlChainSmem& sm = chain_smem[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    if (!f.lvl) {
        // no schedule: every warp takes the next parked flow (dynamic balancing through one atomic counter)
        for (;;) {
            uint32_t i = 0u;
            if (lane == 0) i = atomicAdd(&f.counters[1], 1u);
            i = __shfl_sync(FL_FULL, i, 0);
            if (i >= count) return;
            const uint32_t cur = f.parked[i];
            fl_flow_warp(f, cur, fl_ld_cg(&f.xbuf[cur]), fl_ld_cg(&f.hbuf[cur]), true, sm, f.next_list != nullptr);
        }
    }
    // scheduled by level: every warp looks at 32 parked flows at a time (one per lane), walks those of this level
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t base = warp * 32u; base < count; base += nwarps * 32u) {
        const uint32_t i = base + (uint32_t)lane;
        uint32_t mine = FL_NONE;
        if (i < count) {
            mine = f.parked[i];
            const uint32_t lv = f.lvl[mine];
            if ((lv < f.top_level ? lv : f.top_level) != f.level) mine = FL_NONE;
        }
        uint32_t todo = __ballot_sync(FL_FULL, mine != FL_NONE);
        while (todo) {
            const int src = __ffs((int)todo) - 1;
            todo &= todo - 1u;
            const uint32_t cur = __shfl_sync(FL_FULL, mine, src);
            fl_flow_warp(f, cur, fl_ld_cg(&f.xbuf[cur]), fl_ld_cg(&f.hbuf[cur]), true, sm, f.next_list != nullptr);
        }
    }
#endif
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
    sm.in[lane] = w.t;
    __syncwarp();
    {
        double r = rt_prev;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const double tk = sm.in[k];
            if ((uint32_t)k < nproc) r = 0.0 + (r + tk);
            sm.out[k] = r;
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
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            double zk = sm.in[k];
            const double msk = sm.aux1[k];
            const double dk = sm.aux2[k];
            if ((uint32_t)k < nproc) {
                if (msk == msk) {
                    const double slope = (zk - zp) / dk;
                    if (slope > msk) zk = zp + msk * dk;
                }
                zp = zk;
            }
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
