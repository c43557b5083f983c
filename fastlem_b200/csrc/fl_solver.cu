// fl_solver.cu -- host orchestration of the device-resident generate() loop + the C ABI
// (include/fastlem_b200.h).  Reference path: src/lem/generator.rs:90-213.
//
// Compiled by nvcc for sm_100a (product) or, with -DFL_EMU, by g++ as a serial host emulation used only
// by the CPU test tier (see fl_rt.h).
//
// Two sweep implementations (option "sweep"):
//   0  level-synchronous: nodes sorted by tree depth, one launch per level and sweep (simple; launch bound)
//   1  path-decomposed  : fl_paths.cuh -- the sites are renumbered so that every heavy path is contiguous and
//                         one thread walks a path; ~20 rounds per sweep instead of ~1000 levels
//   2  as 1, and paths of >= FL_LONG_PATH sites are walked by a whole warp (default)
#include "../../include/fastlem_b200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "fl_flood.h"
#include "fl_kernels.cuh"
#include "fl_flow.cuh"
#include "fl_elev.cuh"
#include "fl_floodgpu.cuh"
#include "fl_paths.cuh"

#ifdef FL_EMU
thread_local fl_dim3 threadIdx, blockIdx, blockDim, gridDim;
#endif

namespace {

double wall_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#define FL_MAX_ROUNDS 4096u
#define FL_KEY_BASE 254u  // sort key of a segment head = FL_KEY_BASE - nesting height (one 8-bit radix pass)
#define FL_WARP_LEVEL_MAX 8192u  // K5: levels with at most this many segments run one warp per segment
enum Stage { ST_RECV = 0, ST_LABEL, ST_LAKE, ST_ORDER, ST_AREA, ST_ELEV, ST_COUNT };

// One numbering of the sites and everything stored in it.
struct Layout {
    uint32_t* row_ptr = nullptr;
    uint32_t* col = nullptr;
    double* dist = nullptr;
    uint8_t* rev = nullptr;
    double* areas = nullptr;
    double* erod = nullptr;
    double* uplift = nullptr;
    double* tan = nullptr;  // null = max_slope None everywhere
    uint8_t* is_outlet = nullptr;
    uint32_t* rank = nullptr;     // flood order, valid when ctx.rank_ready
    uint32_t* orig_of = nullptr;  // this numbering -> the caller's
    double* elev = nullptr;
    double* drecv = nullptr;
    uint32_t* recv = nullptr;
    uint32_t* cmask = nullptr;
    uint32_t* lvl = nullptr;  // nesting height of the site's segment in the last K5 (sweep 3)
};

}  // namespace

struct fastlem_ctx {
    int device = 0;
    cudaStream_t stream = 0;
    bool stream_ok = false;
    std::string err;
    std::vector<void**> owned;  // every device allocation, for free_all
    // set_graph's buffers come out of ONE device allocation (cudaMalloc / cudaFree cost milliseconds each on a busy
    // box and there are ~120 buffers): a dry pass adds up the sizes, then the same calls carve the slab
    unsigned char* slab = nullptr;
    size_t slab_size = 0, slab_used = 0, slab_need = 0;
    bool slab_dry = false;
    std::vector<void**> slab_ptrs;

    uint32_t n = 0, nnz = 0;
    bool has_graph = false, has_params = false, has_tan = false;
    // borrowed host pointers (flood order is computed from them on first use)
    std::vector<uint32_t> outlets;

    Layout orig;    // the caller's numbering: static inputs only (+ rank)
    Layout lay[2];  // working numberings (double buffer for renumbering)
    int cur = 0;
    double* d_init = nullptr;  // initial elevations, caller's numbering

    bool rank_ready = false;
    bool layout_valid = false;           // lay[cur] holds a complete numbering (set by reset_layout)
    uint32_t* d_rank_to_node = nullptr;  // current numbering

    // per-iteration state (current numbering)
    unsigned long long* d_pd = nullptr;
    unsigned long long* d_pd2 = nullptr;
    unsigned long long* d_lake_key = nullptr;
    uint32_t* d_label = nullptr;
    uint32_t* d_depth = nullptr;  // also: sort keys
    uint32_t* d_ids = nullptr;
    uint32_t* d_sorted = nullptr;
    uint32_t* d_order = nullptr;  // level order (sweep 0) / sorted path heads (sweep 1)
    uint32_t* d_offs = nullptr;   // n+2
    double* d_A = nullptr;
    double* d_rt = nullptr;
    uint32_t* d_root_of = nullptr;
    // path layout
    uint32_t* d_heavy = nullptr;
    uint32_t* d_plen = nullptr;
    uint32_t* d_len_sorted = nullptr;
    uint32_t* d_hrank = nullptr;
    uint32_t* d_seg_head = nullptr;  // n+1
    uint32_t* d_newpos = nullptr;
    uint32_t* d_deg_new = nullptr;  // n+1
    uint32_t n_levels = 0, n_groups = 0, n_paths = 0;  // groups = (level, long/short) buckets of the path list
    // dataflow sweeps (sweep 3)
    uint32_t* d_state = nullptr;
    double* d_pre = nullptr;
    double* d_post1 = nullptr;
    double* d_post2 = nullptr;
    double* d_xbuf = nullptr;
    double* d_xpost = nullptr;
    uint32_t* d_hbuf = nullptr;
    uint32_t* d_hgt = nullptr;
    uint32_t* d_hpre = nullptr;
    uint32_t* d_iota = nullptr;
    uint32_t* d_parked = nullptr;
    uint32_t* d_nwait = nullptr;
    uint32_t* d_sg_head = nullptr;  // dynamic segments of the dataflow K4: head of each site's segment,
    uint32_t* d_sg_tail = nullptr;  // and per head: tail, waiting sites, published sites
    uint32_t* d_sg_wait = nullptr;
    uint32_t* d_sg_done = nullptr;
    // incremental K4 (fl_flow.cuh): state kept between iterations + per-iteration work lists
    int64_t opt_first_flow = 1;  // the first iteration also uses the dataflow sweeps (layout by subtree sizes)
    // K5 split by nesting height (fl_elev.cuh): queue of run starts for the top of the forest, push masks
    uint2* d_low_list = nullptr;     // K5: (head, receiver) pairs per height below the cut (k_elev_plan -> k_elev_low)
    int low_blocks = 0;              // resident blocks of k_elev_low
    FlQEntry* d_queue = nullptr;   // n entries
    uint32_t* d_pmask = nullptr;   // zero between sweeps (k_elev_top clears what k_elev_plan sets)
    uint32_t push_epoch = 0;
    int top_blocks = 0;            // the most blocks of k_elev_top that can be resident
    // The next iteration's K1 is launched before the host has read this iteration's flags (iterate_flow): it writes into
    // alternate receiver buffers that the next iteration swaps in, so a run that turns out to have ended keeps its state.
    uint32_t* d_recv_alt = nullptr;
    uint32_t* d_cmask_alt = nullptr;
    double* d_drecv_alt = nullptr;
    bool k1_pre = false;           // the alternate buffers hold the next iteration's K1 output
    bool k1_pre_track = false;     // ... and it listed the re-routed sites
    int64_t opt_overlap = 1;       // 0: the host waits for the flags before it launches anything else
    int k1_bulk_blocks = 0;        // resident blocks of k_receivers_bulk (0 = not available)
    int64_t opt_k1_bulk = 1;       // K1 with the CSR stream staged by the bulk-copy engine (fl_paths.cuh)
    int64_t opt_k5_split = 1;      // 0: the round-1 sweep (heads sorted by height, one launch per height)
    int64_t opt_k5_cut = -1;       // cut height: -1 = chosen from the previous iteration's level histogram
    int64_t opt_k5_top_cap = 12288;  // auto cut: at most this many segments go through the queue
    int64_t opt_k5_top_blocks = 0;   // 0 = one block per SM
    uint32_t k5_cut = 0;           // the cut of the next sweep
    bool k5_cut_known = false;
    uint32_t* d_ticket_of = nullptr;  // fused sparse levels of K5
    uint32_t* d_fdone = nullptr;
    uint32_t* d_flvl = nullptr;
    int64_t opt_fuse_levels = 1;
    bool trace_iters = std::getenv("FASTLEM_TRACE_ITERS") != nullptr;
    int64_t opt_key_base = FL_KEY_BASE;  // tests lower it to exercise the deep-nesting path of the head ordering
    uint32_t* d_hsuf = nullptr;
    uint32_t* d_dirty_from = nullptr;
    uint32_t* d_rlist = nullptr;
    uint32_t* d_slist = nullptr;
    uint32_t* d_chg_node = nullptr;
    uint32_t* d_chg_old = nullptr;
    bool k4_valid = false;      // the arrays above and A/pre/post/state/hgt describe the forest of L.recv
    bool k4_last_full = true;   // the previous K4 was a full pass (its counters are not zeroed)
    int64_t opt_flood_device = 1;  // flood order on the device (fl_floodgpu.cuh) when the edge lengths allow it
    int64_t opt_outlet_closed_form = 1;  // the outlets' own ranks by their closed form on the device (0: host replay of the prefix)
    int64_t opt_incremental = 1;
    int64_t opt_incr_div = 16;  // incremental pass when re-routed sites * incr_div <= n
    unsigned long long* d_flow_stats = nullptr;
    unsigned long long* d_tlog = nullptr;  // FL_FLOW_STATS builds only
    uint32_t prev_maxh = 0;
    double* d_tcel = nullptr;
    int sm_count = 148;
    int64_t opt_park_after = 8;
    uint32_t max_degree = 0;
    uint32_t segs_at_rebuild = 0, maxh_at_rebuild = 0;
    bool need_rebuild = true;
    int64_t opt_rebuild_every = 0;  // 0 = adaptive
    // adaptive renumbering: when the segment count grew by `growth` percent, or the nesting height exceeds `height` percent
    // of its value after the last renumbering + 2.  0 = by the number of sites: a renumbering (and the full K4 pass it
    // forces) costs about 0.8 iterations at 1M sites but 2 at 16M, so large models renumber later -- 4 / 150 up to 1M
    // sites, 8 / 250 from 16M on (measured optima, profiles/r2l_rebuild_sweep_*.txt, r2m_*), in between by log2(n)
    int64_t opt_rebuild_height = 0;
    int64_t opt_rebuild_growth = 0;
    // scratch in the caller's numbering (download, debug fetch, kept stages)
    double* d_out_f64 = nullptr;
    uint32_t* d_out_u32 = nullptr;
    uint32_t* d_recv0 = nullptr;
    uint32_t* d_label0 = nullptr;
    bool stages_valid = false;

    uint32_t* d_flags = nullptr;
    uint32_t* h_flags = nullptr;  // pinned
    uint32_t* h_offs = nullptr;   // pinned, n+2 (level offsets of sweeps 0-2)
    uint32_t* h_offs_k = nullptr; // pinned, FL_KEY_BASE+2 (level offsets of sweep 3, indexed by sort key)
    uint32_t* h_rounds = nullptr; // pinned, FL_MAX_ROUNDS+2
    void* d_tmp = nullptr;        // CUB temp storage (sort / scan)
    size_t tmp_bytes = 0;

    int opt_profile = 0;  // 1: stage times (events, no extra synchronisation); 2: + per-kernel times (synchronises)
    bool opt_keep = false;
    int64_t opt_sweep = 3;

    fastlem_stats stats{};
    cudaEvent_t ev[ST_COUNT + 6] = {};  // [9], [10]: K1 of the NEXT iteration; [11]: flags copied
    cudaEvent_t ev_run[2] = {};
    cudaEvent_t ev_k[2] = {};  // "profile"=2: brackets one kernel
};

namespace {

int fail(fastlem_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}

#define FL_CK(expr)                                                                              \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return fail(c, FASTLEM_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)
#define FL_RC(expr)          \
    do {                     \
        int rc__ = (expr);   \
        if (rc__) return rc__; \
    } while (0)

template <class T> cudaError_t dalloc(fastlem_ctx* c, T*& p, size_t count) {
    const size_t bytes = ((count ? count : 1) * sizeof(T) + 255) & ~(size_t)255;
    if (c->slab_dry) { c->slab_need += bytes; return cudaSuccess; }
    if (c->slab && c->slab_used + bytes <= c->slab_size) {
        p = (T*)(c->slab + c->slab_used);
        c->slab_used += bytes;
        c->slab_ptrs.push_back((void**)&p);
        return cudaSuccess;
    }
    if (p) { fl_free(p, c->stream); p = nullptr; }
    void* v = nullptr;
    cudaError_t e = fl_malloc(&v, count * sizeof(T));
    p = (T*)v;
    bool known = false;
    for (void** q : c->owned) known = known || (q == (void**)&p);
    if (!known) c->owned.push_back((void**)&p);
    return e;
}

void free_all(fastlem_ctx* c) {
    for (void** q : c->owned)
        if (*q) { fl_free(*q, c->stream); *q = nullptr; }
    c->owned.clear();
    for (void** q : c->slab_ptrs) *q = nullptr;
    c->slab_ptrs.clear();
    if (c->slab) fl_free(c->slab, c->stream);
    c->slab = nullptr;
    c->slab_size = c->slab_used = c->slab_need = 0;
    if (c->h_offs) fl_free_host(c->h_offs);
    c->h_offs = nullptr;
    c->has_graph = c->has_params = c->rank_ready = c->stages_valid = c->has_tan = c->layout_valid = false;
}

// pinned host buffer for the per-level offsets of the level / path sweeps (allocated on first use: the default
// dataflow sweeps read back FL_KEY_BASE+2 words only)
int ensure_h_offs(fastlem_ctx* c) {
    if (c->h_offs) return FASTLEM_OK;
    void* ho = nullptr;
    FL_CK(fl_malloc_host(&ho, sizeof(uint32_t) * ((size_t)c->n + 2)));
    c->h_offs = (uint32_t*)ho;
    return FASTLEM_OK;
}

inline unsigned blocks_for(uint32_t count, unsigned block = 256) { return (count + block - 1u) / block; }
inline Layout& L_(fastlem_ctx* c) { return c->lay[c->cur]; }

#define LAUNCH_N(kernel, count, ...)                                                  \
    do {                                                                              \
        if ((count) > 0) {                                                            \
            FL_LAUNCH(kernel, blocks_for(count), 256, c->stream, __VA_ARGS__);        \
            c->stats.kernel_launches++;                                               \
        }                                                                             \
    } while (0)

int stage_mark(fastlem_ctx* c, int k) {
    if (c->opt_profile) FL_CK(fl_event_record(c->ev[k], c->stream));
    return FASTLEM_OK;
}

// "profile"=2: device time of single kernels (fastlem_stats.ms_kernel); k_begin before the launch, k_end after it
int k_begin(fastlem_ctx* c) {
    if (c->opt_profile >= 2) FL_CK(fl_event_record(c->ev_k[0], c->stream));
    return FASTLEM_OK;
}
int k_end(fastlem_ctx* c, int id) {
    if (c->opt_profile < 2) return FASTLEM_OK;
    FL_CK(fl_event_record(c->ev_k[1], c->stream));
    FL_CK(fl_event_sync(c->ev_k[1]));
    float ms = 0.f;
    FL_CK(fl_event_elapsed(&ms, c->ev_k[0], c->ev_k[1]));
    c->stats.ms_kernel[id] += ms;
    c->stats.n_kernel[id] += 1;
    return FASTLEM_OK;
}

// the flag words on their way to the host; the caller may launch more work before it waits for them
int read_flags_begin(fastlem_ctx* c) {
    FL_CK(fl_d2h(c->h_flags, c->d_flags, sizeof(uint32_t) * FL_N_FLAGS, c->stream));
    FL_CK(fl_event_record(c->ev[ST_COUNT + 5], c->stream));
    return FASTLEM_OK;
}
int read_flags_end(fastlem_ctx* c) {
    FL_CK(fl_event_sync(c->ev[ST_COUNT + 5]));
    return FASTLEM_OK;
}

int read_flags(fastlem_ctx* c) {
    FL_CK(fl_d2h(c->h_flags, c->d_flags, sizeof(uint32_t) * FL_N_FLAGS, c->stream));
    FL_CK(fl_stream_sync(c->stream));
    return FASTLEM_OK;
}

// pointer jumping on a packed (pointer, distance) table until stable.  Rounds go out in batches of three
// with one flag read-back per batch (the flag is cleared before the batch's last round only, so it tells
// whether that round still changed anything).
int jump_loop(fastlem_ctx* c, unsigned long long* pd) {
    const uint32_t n = c->n;
    for (int batch = 0; batch < 24; ++batch) {
        LAUNCH_N(k_jump, n, n, pd, c->d_flags);
        LAUNCH_N(k_jump, n, n, pd, c->d_flags);
        FL_CK(fl_memset(c->d_flags + FL_FLAG_JUMP, 0, sizeof(uint32_t), c->stream));
        LAUNCH_N(k_jump, n, n, pd, c->d_flags);
        c->stats.n_labels += 3;
        FL_RC(read_flags(c));
        if (!c->h_flags[FL_FLAG_JUMP]) return FASTLEM_OK;
    }
    return fail(c, FASTLEM_E_STATE, "pointer jumping did not converge (cycle in receivers?)");
}

// K2: (root, depth) of every node in the receiver forest -> d_pd
int run_labels(fastlem_ctx* c) {
    LAUNCH_N(k_jump_init, c->n, c->n, L_(c).recv, c->d_pd);
    return jump_loop(c, c->d_pd);
}

// FASTLEM_TRACE=1: wall-clock phase times on stderr (debug aid)
struct FlTrace {
    bool on;
    double t;
    const char* what;
    explicit FlTrace(const char* w) : on(std::getenv("FASTLEM_TRACE") != nullptr), t(wall_ms()), what(w) {}
    void mark(const char* phase) {
        if (!on) return;
        const double now = wall_ms();
        std::fprintf(stderr, "[fastlem trace] %s: %s %.3f ms\n", what, phase, now - t);
        t = now;
    }
};

// device allocations that live for one call
struct FlTmpAlloc {
    std::vector<void*> p;
    unsigned char* slab = nullptr;
    size_t size = 0, used = 0;
    cudaStream_t stream;
    explicit FlTmpAlloc(cudaStream_t s) : stream(s) {}
    ~FlTmpAlloc() { for (void* q : p) fl_free(q, stream); }
    cudaError_t reserve(size_t bytes) {  // one allocation for everything that fits
        void* v = nullptr;
        cudaError_t e = fl_malloc(&v, bytes);
        if (e == cudaSuccess) { p.push_back(v); slab = (unsigned char*)v; size = bytes; used = 0; }
        return e;
    }
    template <class T> cudaError_t get(T*& out, size_t count) {
        const size_t bytes = ((count ? count : 1) * sizeof(T) + 255) & ~(size_t)255;
        if (slab && used + bytes <= size) { out = (T*)(slab + used); used += bytes; return cudaSuccess; }
        void* v = nullptr;
        cudaError_t e = fl_malloc(&v, bytes);
        if (e == cudaSuccess) p.push_back(v);
        out = (T*)v;
        return e;
    }
};

// the graph back on the host, for the exact heap replay (fl_flood.cpp) that inputs with tied edge lengths need
struct HostCsr {
    std::vector<uint32_t> row_ptr, col;
    std::vector<double> dist;
};
int download_csr(fastlem_ctx* c, HostCsr& h) {
    h.row_ptr.resize((size_t)c->n + 1); h.col.resize(c->nnz); h.dist.resize(c->nnz);
    FL_CK(fl_d2h(h.row_ptr.data(), c->orig.row_ptr, sizeof(uint32_t) * ((size_t)c->n + 1), c->stream));
    if (c->nnz) {
        FL_CK(fl_d2h(h.col.data(), c->orig.col, sizeof(uint32_t) * c->nnz, c->stream));
        FL_CK(fl_d2h(h.dist.data(), c->orig.dist, sizeof(double) * c->nnz, c->stream));
    }
    FL_CK(fl_stream_sync(c->stream));
    return FASTLEM_OK;
}

// flood order on the device (fl_floodgpu.cuh).  *done = false: the graph has equal / non-positive edge lengths (or
// rows too long for the reverse-slot table) and the exact host replay must be used instead.
int device_flood_rank(fastlem_ctx* c, bool* done) {
    *done = false;
    FL_RANGE("fastlem flood order (device)");
    const uint32_t n = c->n, nnz = c->nnz;
    if (c->max_degree >= 255u || c->outlets.empty() || nnz == 0u) return FASTLEM_OK;
    FlTrace tr("flood");
    FlTmpAlloc tmp(c->stream);
    const size_t n1 = (size_t)n + 1;
    // everything below comes out of one allocation: 17 bytes per slot (tree flags, two key arrays) and ~150 per site
    FL_CK(tmp.reserve((size_t)nnz * 17 + n1 * 160 + ((size_t)1 << 20) + 2 * c->tmp_bytes));
    FlFloodG g;
    g.n = n; g.src = c->outlets[0];
    g.row_ptr = c->orig.row_ptr; g.col = c->orig.col; g.dist = c->orig.dist; g.rev = c->orig.rev;
    g.is_outlet = c->orig.is_outlet;
    uint32_t *frontier_a = nullptr, *frontier_b = nullptr, *ids_a = nullptr, *ids_b = nullptr, *gstart = nullptr,
             *gscan = nullptr, *outlet_rank = nullptr;
    unsigned long long *keys_a = nullptr, *keys_b = nullptr, *k64a = nullptr, *k64b = nullptr, *szs = nullptr,
                       *scan = nullptr, *pd = nullptr;
    FL_CK(tmp.get(g.comp, n)); FL_CK(tmp.get(g.link, n)); FL_CK(tmp.get(g.best, n)); FL_CK(tmp.get(g.pick, n));
    FL_CK(tmp.get(g.mst, nnz)); FL_CK(tmp.get(g.par, n1)); FL_CK(tmp.get(g.wbits, n1)); FL_CK(tmp.get(g.nga, n1));
    FL_CK(tmp.get(g.cnt, n1)); FL_CK(tmp.get(g.size, n1)); FL_CK(tmp.get(g.flags, 8));
    FL_CK(tmp.get(keys_a, nnz)); FL_CK(tmp.get(keys_b, nnz));
    FL_CK(fl_memset(g.flags, 0, sizeof(uint32_t) * 8, c->stream));
    size_t need = 0, t1 = 0;
    FL_CK(fl_sort_keys64(nullptr, need, keys_a, keys_b, nnz, c->stream, true));
    FL_CK(fl_sort_pairs64(nullptr, t1, k64a, k64b, ids_a, ids_b, n, 64, c->stream, true));
    if (t1 > need) need = t1;
    FL_CK(fl_exclusive_sum64(nullptr, t1, szs, scan, n, c->stream, true));
    if (t1 > need) need = t1;
    if (c->tmp_bytes > need) need = c->tmp_bytes;
    unsigned char* cub_raw = nullptr;
    FL_CK(tmp.get(cub_raw, need));
    void* const cub_tmp = cub_raw;

    tr.mark("alloc");
    // 1. all edge lengths positive, finite and pairwise distinct?
    LAUNCH_N(k_flg_edge_keys, n, g, keys_a);
    FL_CK(fl_sort_keys64(cub_tmp, need, keys_a, keys_b, nnz, c->stream, false));
    LAUNCH_N(k_flg_dup_check, nnz, nnz, keys_b, g.flags);
    uint32_t hf[8];
    FL_CK(fl_d2h(hf, g.flags, sizeof(hf), c->stream));
    FL_CK(fl_stream_sync(c->stream));
    if (hf[3]) return FASTLEM_OK;  // ties: the host replay reproduces the heap's behaviour

    tr.mark("distinct check");
    // 2. the outlets' own ranks (equal keys 0.0: heap behaviour): closed form + validity check on the device
    //    (fl_floodgpu.cuh, k_flg_outlet_*); the exact host replay of that prefix only when the check fails
    const uint32_t n_out = (uint32_t)c->outlets.size();  // distinct (fastlem_set_parameters rejects duplicates) = |S|
    {
        uint32_t *d_outlets = nullptr, *pushes = nullptr, *before = nullptr;
        FL_CK(tmp.get(outlet_rank, n)); FL_CK(tmp.get(d_outlets, n_out)); FL_CK(tmp.get(pushes, n_out));
        FL_CK(tmp.get(before, n_out));
        FL_CK(fl_h2d(d_outlets, c->outlets.data(), sizeof(uint32_t) * n_out, c->stream));
        FL_CK(fl_memset(outlet_rank, 0xFF, sizeof(uint32_t) * n, c->stream));
        bool closed = c->opt_outlet_closed_form != 0 && !hf[7];
        if (closed) {
            const FlOutletPrefix f = flg_outlet_prefix(n_out);
            LAUNCH_N(k_flg_outlet_rank, n_out, f, d_outlets, outlet_rank);
            LAUNCH_N(k_flg_outlet_pushes, n_out, g, n_out, d_outlets, outlet_rank, pushes);
            FL_CK(fl_exclusive_sum(cub_tmp, need, pushes, before, n_out, c->stream, false));
            LAUNCH_N(k_flg_outlet_check, n_out, n_out, before, g.flags);
            FL_CK(fl_d2h(hf, g.flags, sizeof(hf), c->stream));
            FL_CK(fl_stream_sync(c->stream));
            closed = !hf[7];
        }
        c->stats.outlet_ranks_on_device = closed ? 1u : 0u;
        if (!closed) {
            HostCsr h;
            FL_RC(download_csr(c, h));
            std::vector<uint32_t> orank(n);
            fl_flood_rank_prefix(n, h.row_ptr.data(), h.col.data(), h.dist.data(), c->outlets.data(), n_out, orank.data(),
                                 n_out);
            FL_CK(fl_h2d(outlet_rank, orank.data(), sizeof(uint32_t) * n, c->stream));
            FL_CK(fl_stream_sync(c->stream));
        }
    }

    tr.mark("outlet prefix");
    // 3. Boruvka: minimum spanning forest of the graph with the outlets contracted
    LAUNCH_N(k_flg_init, n + 1, g);
    FL_CK(fl_memset(g.mst, 0, nnz, c->stream));
    for (int round = 0; round < 64; ++round) {
        FL_CK(fl_memset(g.flags, 0, sizeof(uint32_t), c->stream));
        LAUNCH_N(k_flg_min, n, g);
        LAUNCH_N(k_flg_pick, n, g);
        LAUNCH_N(k_flg_hook, n, g);
        LAUNCH_N(k_flg_relabel, n, g);
        LAUNCH_N(k_flg_reset, n, g);
        FL_CK(fl_d2h(hf, g.flags, sizeof(uint32_t), c->stream));
        FL_CK(fl_stream_sync(c->stream));
        if (!hf[0]) break;
        if (round == 63) return fail(c, FASTLEM_E_STATE, "flood order: spanning forest did not converge");
    }

    tr.mark("boruvka");
    // 4. root the tree at the source (one CTA walks all levels over a compact copy of the tree's adjacency)
    FL_CK(tmp.get(frontier_a, n1)); FL_CK(tmp.get(frontier_b, n1));
    FL_CK(tmp.get(g.tptr, n1)); FL_CK(tmp.get(g.tcol, 2 * (size_t)n)); FL_CK(tmp.get(g.twb, 2 * (size_t)n));
    LAUNCH_N(k_flg_tree_degree, n + 1, g, frontier_b);
    FL_CK(fl_exclusive_sum(cub_tmp, need, frontier_b, g.tptr, n + 1, c->stream, false));
    LAUNCH_N(k_flg_tree_fill, n, g);
    uint32_t* cnt3 = g.flags + 4;
    FL_CK(fl_memset(cnt3, 0, sizeof(uint32_t) * 3, c->stream));
    LAUNCH_N(k_flg_root_init, n, g, frontier_a, cnt3);
    FL_LAUNCH(k_flg_root_walk, 1, 1024, c->stream, g, frontier_a, frontier_b, cnt3);
    c->stats.kernel_launches++;
    FL_CK(fl_stream_sync(c->stream));
    tr.mark("rooting");
    // 5. nearest heavier ancestor, sizes of the nga subtrees
    for (int pass = 0; pass < 64; ++pass) {
        FL_CK(fl_memset(g.flags + 2, 0, sizeof(uint32_t), c->stream));
        LAUNCH_N(k_flg_nga, n, g, 1u << 16);
        FL_CK(fl_d2h(hf, g.flags, sizeof(uint32_t) * 4, c->stream));
        FL_CK(fl_stream_sync(c->stream));
        if (!hf[2]) break;
        if (pass == 63) return fail(c, FASTLEM_E_STATE, "flood order: ancestor search did not converge");
    }
    LAUNCH_N(k_flg_count_children, n, g);
    LAUNCH_N(k_flg_bias_counts, n + 1, g);
    LAUNCH_N(k_flg_sizes, n, g);

    FL_CK(fl_stream_sync(c->stream));
    tr.mark("nga + sizes");
    // 6. siblings by (nga parent, parent edge length); lighter siblings' sizes by a scan
    FL_CK(tmp.get(k64a, n)); FL_CK(tmp.get(k64b, n)); FL_CK(tmp.get(ids_a, n)); FL_CK(tmp.get(ids_b, n));
    FL_CK(tmp.get(szs, n)); FL_CK(tmp.get(scan, n)); FL_CK(tmp.get(gstart, n)); FL_CK(tmp.get(gscan, n));
    FL_CK(tmp.get(pd, n1));
    LAUNCH_N(k_flg_sort_keys, n, g, k64a, ids_a);
    FL_CK(fl_sort_pairs64(cub_tmp, need, k64a, k64b, ids_a, ids_b, n, 64, c->stream, false));
    LAUNCH_N(k_flg_parent_keys, n, g, ids_b, k64a);
    FL_CK(fl_sort_pairs64(cub_tmp, need, k64a, k64b, ids_b, ids_a, n, 33, c->stream, false));  // stable
    LAUNCH_N(k_flg_group, n, g, k64b, ids_a, szs, gstart);
    FL_CK(fl_exclusive_sum64(cub_tmp, need, szs, scan, n, c->stream, false));
    FL_CK(fl_inclusive_max(cub_tmp, need, gstart, gscan, n, c->stream, false));
    LAUNCH_N(k_flg_terms, n + 1, g, n_out, k64b, ids_a, scan, gscan, pd);

    // 7. T(v) = sum of the terms along the nga pointers
    for (int batch = 0; batch < 24; ++batch) {
        LAUNCH_N(k_jump, n + 1, n + 1, pd, c->d_flags);
        LAUNCH_N(k_jump, n + 1, n + 1, pd, c->d_flags);
        FL_CK(fl_memset(c->d_flags + FL_FLAG_JUMP, 0, sizeof(uint32_t), c->stream));
        LAUNCH_N(k_jump, n + 1, n + 1, pd, c->d_flags);
        FL_RC(read_flags(c));
        if (!c->h_flags[FL_FLAG_JUMP]) break;
        if (batch == 23) return fail(c, FASTLEM_E_STATE, "flood order: final sums did not converge");
    }
    LAUNCH_N(k_flg_finish, n, g, pd, outlet_rank, c->orig.rank);
#ifdef FL_EMU
    if (const char* path = std::getenv("FL_FLOOD_DUMP")) {  // host emulation only: intermediate arrays for debugging
        if (FILE* fp = std::fopen(path, "wb")) {
            std::fwrite(g.par, 4, n1, fp); std::fwrite(g.wbits, 8, n1, fp); std::fwrite(g.nga, 4, n1, fp);
            std::fwrite(g.size, 4, n1, fp); std::fwrite(pd, 8, n1, fp);
            std::fclose(fp);
        }
    }
#endif
    FL_CK(fl_stream_sync(c->stream));
    FL_CK(fl_last_error());
    tr.mark("sort + scan + jump");
    *done = true;
    return FASTLEM_OK;
}

// flood order: computed once per (graph, outlets) in the caller's numbering -- on the device when the edge lengths
// are distinct (fl_floodgpu.cuh), else by the exact host replay (fl_flood.cpp)
int ensure_rank(fastlem_ctx* c) {
    if (c->rank_ready) return FASTLEM_OK;
    double t0 = wall_ms();
    const uint32_t n = c->n;
    bool done = false;
    const uint64_t launches_before = c->stats.kernel_launches;
    if (c->opt_flood_device) FL_RC(device_flood_rank(c, &done));
    c->stats.flood_on_device = done ? 1u : 0u;
    if (!done) {
        c->stats.outlet_ranks_on_device = 0u;
        c->stats.kernel_launches = launches_before;
        HostCsr h;  // (the graph lives in HBM; the caller's arrays are not kept)
        FL_RC(download_csr(c, h));
        std::vector<uint32_t> rank(n);
        fl_flood_rank(n, h.row_ptr.data(), h.col.data(), h.dist.data(), c->outlets.data(), (uint32_t)c->outlets.size(),
                      rank.data());
        FL_CK(fl_h2d(c->orig.rank, rank.data(), sizeof(uint32_t) * n, c->stream));
        FL_CK(fl_stream_sync(c->stream));
    }
    if (c->layout_valid) {  // into the current numbering (otherwise reset_layout does it at the next run)
        Layout& L = L_(c);
        LAUNCH_N(k_gather_u32, n, n, L.orig_of, c->orig.rank, L.rank);
        FL_CK(fl_memset(c->d_rank_to_node, 0xFF, sizeof(uint32_t) * n, c->stream));
        LAUNCH_N(k_rank_inverse, n, n, L.rank, c->d_rank_to_node);
    }
    c->rank_ready = true;
    c->stats.ms_flood_rank = wall_ms() - t0;
    return FASTLEM_OK;
}

// K3: connect every lake basin and reverse its in-basin path (labels must be in d_pd)
int run_lakes(fastlem_ctx* c) {
    FL_RANGE("fastlem K3 lake removal");
    const uint32_t n = c->n;
    Layout& L = L_(c);
    FL_RC(ensure_rank(c));
    LAUNCH_N(k_labels_only, n, n, c->d_pd, c->d_label);
    if (c->opt_keep) {
        LAUNCH_N(k_unpermute_ids, n, n, L.orig_of, L.recv, c->d_recv0);
        LAUNCH_N(k_unpermute_ids, n, n, L.orig_of, c->d_label, c->d_label0);
        c->stages_valid = true;
    }
    FL_CK(fl_memset(c->d_lake_key, 0xFF, sizeof(unsigned long long) * n, c->stream));
    LAUNCH_N(k_lake_min, n, n, L.row_ptr, L.col, c->d_label, L.is_outlet, L.rank, c->d_lake_key);
    LAUNCH_N(k_lake_reverse, n, n, L.row_ptr, L.col, L.dist, L.is_outlet, c->d_rank_to_node, c->d_lake_key, c->d_label,
             L.recv, L.drecv);
    c->stats.n_lakes += 3;
    c->stats.lake_iterations++;
    return FASTLEM_OK;
}

int profile_accumulate(fastlem_ctx* c) {
    if (!c->opt_profile) return FASTLEM_OK;
    double* acc[ST_COUNT] = {&c->stats.ms_receivers, &c->stats.ms_labels, &c->stats.ms_lakes,
                             &c->stats.ms_order,     &c->stats.ms_area,   &c->stats.ms_elevation};
    for (int k = 0; k < ST_COUNT; ++k) {
        float ms = 0.f;
        FL_CK(fl_event_elapsed(&ms, c->ev[k], c->ev[k + 1]));
        *acc[k] += ms;
    }
    return FASTLEM_OK;
}

// ------------------------------------------------------------------------------------------------
// sweep 0: one loop body of generator.rs:140-210, level-synchronous
// ------------------------------------------------------------------------------------------------
int iterate_levels(fastlem_ctx* c, bool first, bool* changed_out) {
    const uint32_t n = c->n;
    Layout& L = L_(c);
    c->k4_valid = false;
    FL_RC(ensure_h_offs(c));
    FL_CK(fl_memset(c->d_flags, 0, sizeof(uint32_t) * FL_N_FLAGS, c->stream));
    FL_RC(stage_mark(c, 0));

    LAUNCH_N(k_receivers, n, n, L.row_ptr, L.col, L.dist, L.elev, L.is_outlet, L.recv, L.drecv, c->d_flags);
    c->stats.n_receivers++;
    FL_RC(stage_mark(c, 1));

    FL_RC(run_labels(c));
    const bool has_lake = c->h_flags[FL_FLAG_LAKE] != 0;
    FL_RC(stage_mark(c, 2));

    c->stages_valid = false;
    if (has_lake) {
        FL_RC(run_lakes(c));
        FL_RC(run_labels(c));
    }
    FL_RC(stage_mark(c, 3));

    // ordering: sort nodes by depth
    LAUNCH_N(k_labels_finalize, n, n, c->d_pd, L.is_outlet, L.areas, c->d_label, c->d_depth, c->d_ids, c->d_A, c->d_rt);
    FL_CK(fl_sort_pairs(c->d_tmp, c->tmp_bytes, c->d_depth, c->d_sorted, c->d_ids, c->d_order, n, 32, c->stream, false));
    FL_CK(fl_memset(c->d_flags + FL_FLAG_MAXDEPTH, 0xFF, sizeof(uint32_t), c->stream));
    LAUNCH_N(k_level_offsets, n, n, c->d_sorted, c->d_offs, c->d_flags);
    c->stats.n_order += 3;
    FL_RC(read_flags(c));
    const uint32_t maxd = c->h_flags[FL_FLAG_MAXDEPTH];
    FL_RC(stage_mark(c, 4));
    if (maxd != FL_NONE) {
        FL_CK(fl_d2h(c->h_offs, c->d_offs, sizeof(uint32_t) * ((size_t)maxd + 2), c->stream));
        FL_CK(fl_stream_sync(c->stream));
        if (first) c->stats.depth_first = maxd + 1;
        c->stats.depth_last = maxd + 1;
        for (uint32_t lv = maxd + 1; lv-- > 0;) {  // K4: deepest level first
            const uint32_t b = c->h_offs[lv], cnt = c->h_offs[lv + 1] - b;
            LAUNCH_N(k_area_level, cnt, b, cnt, c->d_order, L.row_ptr, L.col, L.recv, L.areas, c->d_A);
        }
        c->stats.n_area += maxd + 1;
        FL_RC(stage_mark(c, 5));
        for (uint32_t lv = 0; lv <= maxd; ++lv) {  // K5: roots first
            const uint32_t b = c->h_offs[lv], cnt = c->h_offs[lv + 1] - b;
            LAUNCH_N(k_elev_level, cnt, b, cnt, (int)lv, c->d_order, L.recv, c->d_label, L.drecv, c->d_A, L.erod,
                     L.uplift, c->has_tan ? L.tan : nullptr, L.elev, c->d_rt, c->d_flags);
        }
        c->stats.n_elevation += maxd + 1;
    } else {
        FL_RC(stage_mark(c, 5));
    }
    FL_RC(stage_mark(c, 6));
    FL_RC(read_flags(c));
    FL_CK(fl_last_error());
    *changed_out = c->h_flags[FL_FLAG_CHANGED] != 0;
    return profile_accumulate(c);
}

// ------------------------------------------------------------------------------------------------
// sweep 1: renumber the sites so that the heavy paths of the CURRENT receiver forest are contiguous
// (fl_paths.cuh).  `weight` ranks the children of a node (previous drainage areas).
// ------------------------------------------------------------------------------------------------
int rebuild_layout(fastlem_ctx* c, const double* weight) {
    const uint32_t n = c->n;
    FL_RC(ensure_h_offs(c));
    Layout& L = c->lay[c->cur];
    Layout& M = c->lay[c->cur ^ 1];

    LAUNCH_N(k_heavy, n, n, L.row_ptr, L.col, L.recv, L.cmask, weight, c->d_heavy);
    LAUNCH_N(k_chain_init, n, n, L.recv, c->d_heavy, c->d_pd);
    FL_RC(jump_loop(c, c->d_pd));  // -> (path head, position in path)
    LAUNCH_N(k_path_len, n, n, c->d_heavy, c->d_pd, c->d_plen);
    LAUNCH_N(k_nest_init, n, n, L.recv, c->d_pd, c->d_pd2);
    FL_RC(jump_loop(c, c->d_pd2));  // -> (root path head, nesting level)
    FL_CK(fl_memset(c->d_flags + FL_FLAG_MAXDEPTH, 0, sizeof(uint32_t), c->stream));
    const int split = c->opt_sweep >= 2 ? 1 : 0;
    LAUNCH_N(k_path_keys, n, n, c->d_pd, c->d_pd2, c->d_plen, split, c->d_depth, c->d_ids, c->d_flags);
    FL_RC(read_flags(c));
    const uint32_t max_key = c->h_flags[FL_FLAG_MAXDEPTH];
    int bits = 1;
    while (bits < 32 && (1ull << bits) <= (unsigned long long)max_key + 1ull) ++bits;  // FL_NONE sorts last
    FL_CK(fl_sort_pairs(c->d_tmp, c->tmp_bytes, c->d_depth, c->d_sorted, c->d_ids, c->d_order, n, bits, c->stream,
                        false));
    FL_CK(fl_memset(c->d_flags + FL_FLAG_MAXDEPTH, 0xFF, sizeof(uint32_t), c->stream));
    FL_CK(fl_memset(c->d_offs, 0xFF, sizeof(uint32_t) * ((size_t)max_key + 2), c->stream));
    LAUNCH_N(k_level_offsets, n, n, c->d_sorted, c->d_offs, c->d_flags);
    FL_RC(read_flags(c));
    if (c->h_flags[FL_FLAG_MAXDEPTH] != max_key) return fail(c, FASTLEM_E_STATE, "layout: level bookkeeping broke");
    c->n_groups = max_key + 1;
    c->n_levels = split ? (max_key / 2 + 1) : (max_key + 1);
    c->n_paths = c->h_flags[FL_FLAG_REACHED];
    FL_CK(fl_d2h(c->h_offs, c->d_offs, sizeof(uint32_t) * ((size_t)c->n_groups + 1), c->stream));

    LAUNCH_N(k_path_gather, c->n_paths, c->n_paths, c->d_order, c->d_plen, c->d_len_sorted, c->d_hrank);
    FL_CK(fl_exclusive_sum(c->d_tmp, c->tmp_bytes, c->d_len_sorted, c->d_seg_head, c->n_paths, c->stream, false));
    LAUNCH_N(k_newpos, n, n, c->d_pd, c->d_hrank, c->d_seg_head, c->d_newpos);

    // renumber everything: L -> M
    FL_CK(fl_memset(c->d_deg_new + n, 0, sizeof(uint32_t), c->stream));
    LAUNCH_N(k_deg_scatter, n, n, L.row_ptr, c->d_newpos, c->d_deg_new);
    FL_CK(fl_exclusive_sum(c->d_tmp, c->tmp_bytes, c->d_deg_new, M.row_ptr, n + 1, c->stream, false));
    LAUNCH_N(k_permute_rows, n, n, L.row_ptr, L.col, L.dist, L.rev, c->d_newpos, M.row_ptr, M.col, M.dist, M.rev);
    FlNodeArrays a;
    a.areas = L.areas; a.erod = L.erod; a.uplift = L.uplift; a.tan = c->has_tan ? L.tan : nullptr;
    a.elev = L.elev; a.drecv = L.drecv;
    a.areas_n = M.areas; a.erod_n = M.erod; a.uplift_n = M.uplift; a.tan_n = M.tan; a.elev_n = M.elev;
    a.drecv_n = M.drecv;
    a.recv = L.recv; a.cmask = L.cmask; a.rank = c->rank_ready ? L.rank : nullptr; a.orig_of = L.orig_of;
    a.recv_n = M.recv; a.cmask_n = M.cmask; a.rank_n = M.rank; a.orig_of_n = M.orig_of;
    a.is_outlet = L.is_outlet; a.is_outlet_n = M.is_outlet;
    a.lvl = nullptr; a.lvl_n = nullptr;
    LAUNCH_N(k_permute_nodes, n, n, c->d_newpos, a);
    c->cur ^= 1;
    if (c->rank_ready) LAUNCH_N(k_rank_inverse, n, n, M.rank, c->d_rank_to_node);
    FL_CK(fl_stream_sync(c->stream));  // h_offs
    for (uint32_t g = c->n_groups; g-- > 0;)  // keys that do not occur (e.g. a level without long paths)
        if (c->h_offs[g] == FL_NONE) c->h_offs[g] = c->h_offs[g + 1];
    c->stats.n_order += 14;
    c->stats.rebuilds++;
    c->stats.path_levels = c->n_levels;
    c->stats.paths = c->n_paths;
    return FASTLEM_OK;
}

int iterate_paths(fastlem_ctx* c, bool* changed_out) {
    const uint32_t n = c->n;
    c->k4_valid = false;
    FL_CK(fl_memset(c->d_flags, 0, sizeof(uint32_t) * FL_N_FLAGS, c->stream));
    FL_CK(fl_memset(L_(c).cmask, 0, sizeof(uint32_t) * n, c->stream));
    FL_RC(stage_mark(c, 0));
    {
        Layout& L = L_(c);
        LAUNCH_N(k_receivers_mask, n, n, L.row_ptr, L.col, L.dist, L.rev, L.elev, L.is_outlet, L.recv, L.recv, L.drecv,
                 L.cmask, c->d_flags, (uint32_t*)nullptr, (uint32_t*)nullptr);
        c->stats.n_receivers++;
    }
    FL_RC(stage_mark(c, 1));
    FL_RC(read_flags(c));
    const bool has_lake = c->h_flags[FL_FLAG_LAKE] != 0;
    FL_RC(stage_mark(c, 2));
    c->stages_valid = false;
    if (has_lake) {
        Layout& L = L_(c);
        FL_RC(run_labels(c));
        FL_RC(run_lakes(c));
        FL_CK(fl_memset(L.cmask, 0, sizeof(uint32_t) * n, c->stream));
        LAUNCH_N(k_childmask, n, n, L.row_ptr, L.col, L.rev, L.recv, L.cmask);
    }
    FL_RC(stage_mark(c, 3));

    FL_RC(rebuild_layout(c, c->d_A));
    FL_RC(stage_mark(c, 4));

    Layout& L = L_(c);
    const bool split = c->opt_sweep >= 2;
    uint32_t launched = 0;
    for (uint32_t g = c->n_groups; g-- > 0;) {  // K4: innermost paths first
        const uint32_t b = c->h_offs[g], cnt = c->h_offs[g + 1] - b;
        if (!cnt) continue;
        ++launched;
#ifndef FL_EMU
        if (split && (g & 1u) == 0) {
            FL_LAUNCH(k_area_paths_warp, blocks_for(cnt * 32u, 128), 128, c->stream, b, cnt, c->d_seg_head,
                      c->d_len_sorted, L.row_ptr, L.col, L.recv, L.cmask, L.areas, c->d_A);
            continue;
        }
#endif
        FL_LAUNCH(k_area_paths, blocks_for(cnt, 128), 128, c->stream, b, cnt, c->d_seg_head, c->d_len_sorted, L.row_ptr,
                  L.col, L.recv, L.cmask, L.areas, c->d_A);
    }
    c->stats.kernel_launches += launched; c->stats.n_area += launched;
    FL_RC(stage_mark(c, 5));
    launched = 0;
    for (uint32_t g = 0; g < c->n_groups; ++g) {  // K5: root paths first
        const uint32_t b = c->h_offs[g], cnt = c->h_offs[g + 1] - b;
        if (!cnt) continue;
        ++launched;
#ifndef FL_EMU
        if (split && (g & 1u) == 0) {
            FL_LAUNCH(k_elev_paths_warp, blocks_for(cnt * 32u, 128), 128, c->stream, b, cnt, c->d_seg_head,
                      c->d_len_sorted, L.recv, L.drecv, c->d_A, L.erod, L.uplift, c->has_tan ? L.tan : nullptr,
                      L.is_outlet, L.elev, c->d_rt, c->d_root_of, c->d_flags);
            continue;
        }
#endif
        FL_LAUNCH(k_elev_paths, blocks_for(cnt, 128), 128, c->stream, b, cnt, c->d_seg_head, c->d_len_sorted, L.recv,
                  L.drecv, c->d_A, L.erod, L.uplift, c->has_tan ? L.tan : nullptr, L.is_outlet, L.elev, c->d_rt,
                  c->d_root_of, c->d_flags);
    }
    c->stats.kernel_launches += launched; c->stats.n_elevation += launched;
    FL_RC(stage_mark(c, 6));
    FL_RC(read_flags(c));
    FL_CK(fl_last_error());
    *changed_out = c->h_flags[FL_FLAG_CHANGED] != 0;
    return profile_accumulate(c);
}

// ------------------------------------------------------------------------------------------------
// sweep 3: dataflow sweeps on dynamic segments (fl_paths.cuh).  The numbering is rebuilt only when it
// has degraded; segments are whatever chains are contiguous in the current numbering.
// ------------------------------------------------------------------------------------------------
int rebuild_layout_flow(fastlem_ctx* c, const double* weight) {
    FL_RANGE("fastlem renumber sites");
    const uint32_t n = c->n;
    Layout& L = c->lay[c->cur];
    Layout& M = c->lay[c->cur ^ 1];
    LAUNCH_N(k_heavy, n, n, L.row_ptr, L.col, L.recv, L.cmask, weight, c->d_heavy);
    LAUNCH_N(k_chain_init, n, n, L.recv, c->d_heavy, c->d_pd);
    FL_RC(jump_loop(c, c->d_pd));  // -> (path head, position in path)
    LAUNCH_N(k_path_len, n, n, c->d_heavy, c->d_pd, c->d_plen);
    // paths keep the order of their heads; start = exclusive scan of the lengths stored at the heads
    LAUNCH_N(k_head_lengths, n, n, c->d_pd, c->d_plen, c->d_len_sorted);
    FL_CK(fl_exclusive_sum(c->d_tmp, c->tmp_bytes, c->d_len_sorted, c->d_seg_head, n, c->stream, false));
    LAUNCH_N(k_newpos_direct, n, n, c->d_pd, c->d_seg_head, c->d_newpos);
    FL_CK(fl_memset(c->d_deg_new + n, 0, sizeof(uint32_t), c->stream));
    LAUNCH_N(k_deg_scatter, n, n, L.row_ptr, c->d_newpos, c->d_deg_new);
    FL_CK(fl_exclusive_sum(c->d_tmp, c->tmp_bytes, c->d_deg_new, M.row_ptr, n + 1, c->stream, false));
    LAUNCH_N(k_permute_rows, n, n, L.row_ptr, L.col, L.dist, L.rev, c->d_newpos, M.row_ptr, M.col, M.dist, M.rev);
    FlNodeArrays a;
    a.areas = L.areas; a.erod = L.erod; a.uplift = L.uplift; a.tan = c->has_tan ? L.tan : nullptr;
    a.elev = L.elev; a.drecv = L.drecv;
    a.areas_n = M.areas; a.erod_n = M.erod; a.uplift_n = M.uplift; a.tan_n = M.tan; a.elev_n = M.elev;
    a.drecv_n = M.drecv;
    a.recv = L.recv; a.cmask = L.cmask; a.rank = c->rank_ready ? L.rank : nullptr; a.orig_of = L.orig_of;
    a.recv_n = M.recv; a.cmask_n = M.cmask; a.rank_n = M.rank; a.orig_of_n = M.orig_of;
    a.is_outlet = L.is_outlet; a.is_outlet_n = M.is_outlet;
    a.lvl = L.lvl; a.lvl_n = M.lvl;
    LAUNCH_N(k_permute_nodes, n, n, c->d_newpos, a);
    c->cur ^= 1;
    if (c->rank_ready) LAUNCH_N(k_rank_inverse, n, n, M.rank, c->d_rank_to_node);
    c->stats.n_order += 10;
    c->stats.rebuilds++;
    return FASTLEM_OK;
}

// thresholds of the adaptive renumbering (see fastlem_ctx::opt_rebuild_height)
static inline double rebuild_scale(uint32_t n) {
    double t = n > 1000000u ? std::log2((double)n / 1.0e6) / 4.0 : 0.0;
    return t > 1.0 ? 1.0 : t;
}
static inline unsigned long long rebuild_growth_of(const fastlem_ctx* c) {
    return c->opt_rebuild_growth > 0 ? (unsigned long long)c->opt_rebuild_growth
                                     : (unsigned long long)(4.0 + 4.0 * rebuild_scale(c->n) + 0.5);
}
static inline unsigned long long rebuild_height_of(const fastlem_ctx* c) {
    return c->opt_rebuild_height > 0 ? (unsigned long long)c->opt_rebuild_height
                                     : (unsigned long long)(150.0 + 100.0 * rebuild_scale(c->n) + 0.5);
}

// K1 of the current layout: receivers from L.elev into (recv, drecv, cmask); `track`: lists the sites whose receiver
// differs from recv_prev (the previous iteration's final receivers)
int launch_receivers(fastlem_ctx* c, bool track, const uint32_t* recv_prev, uint32_t* recv, double* drecv, uint32_t* cmask) {
    const uint32_t n = c->n;
    Layout& L = L_(c);
    FL_RC(k_begin(c));
    bool bulk = false;
#ifndef FL_EMU
    if (c->opt_k1_bulk && c->k1_bulk_blocks > 0 && n > 0u) {
        // persistent CTAs over tiles of FL_K1B_ROWS rows, the CSR span of a tile bulk-copied one tile ahead
        const uint32_t tiles = (uint32_t)(((unsigned long long)n + FL_K1B_ROWS - 1u) / FL_K1B_ROWS);
        const uint32_t blocks = tiles < (uint32_t)c->k1_bulk_blocks ? tiles : (uint32_t)c->k1_bulk_blocks;
        k_receivers_bulk<<<blocks, 256, FL_K1B_SMEM, c->stream>>>(
            n, tiles, L.row_ptr, L.col, L.dist, L.rev, L.elev, L.is_outlet, recv_prev, recv, drecv, cmask, c->d_flags,
            track ? c->d_chg_node : (uint32_t*)nullptr, track ? c->d_chg_old : (uint32_t*)nullptr);
        c->stats.kernel_launches++;
        bulk = true;
    }
#endif
    if (!bulk)
        LAUNCH_N(k_receivers_mask, n, n, L.row_ptr, L.col, L.dist, L.rev, L.elev, L.is_outlet, recv_prev, recv, drecv, cmask,
                 c->d_flags, track ? c->d_chg_node : (uint32_t*)nullptr, track ? c->d_chg_old : (uint32_t*)nullptr);
    FL_RC(k_end(c, FASTLEM_K_RECEIVERS));
    c->stats.n_receivers++;
    return FASTLEM_OK;
}

// `may_continue`: another iteration may follow this one (the iteration cap has not been reached)
int iterate_flow(fastlem_ctx* c, uint32_t it, bool may_continue, bool* changed_out) {
    FL_RANGE("fastlem iteration");
    const uint32_t n = c->n;
    bool track;
    if (c->k1_pre) {
        // K1 ran at the end of the previous iteration (below), into the alternate buffers: swap them in
        Layout& L = L_(c);
        std::swap(L.recv, c->d_recv_alt);
        std::swap(L.drecv, c->d_drecv_alt);
        std::swap(L.cmask, c->d_cmask_alt);
        std::swap(c->ev[0], c->ev[ST_COUNT + 3]);
        std::swap(c->ev[1], c->ev[ST_COUNT + 4]);
        c->k1_pre = false;
        track = c->k1_pre_track && !c->need_rebuild;
    } else {
        FL_CK(fl_memset(c->d_flags, 0, sizeof(uint32_t) * FL_N_FLAGS, c->stream));
        FL_CK(fl_memset(L_(c).cmask, 0, sizeof(uint32_t) * n, c->stream));
        FL_RC(stage_mark(c, 0));
        // K1 lists the re-routed sites when the K4 state of the previous iteration can be reused
        track = c->opt_incremental != 0 && c->k4_valid && !c->need_rebuild && c->opt_rebuild_every != 1;
        Layout& L = L_(c);
        FL_RC(launch_receivers(c, track, L.recv, L.recv, L.drecv, L.cmask));
        FL_RC(stage_mark(c, 1));
    }
    // While the flags travel to the host: the segment keys of the forest K1 left (k_seg_keys + scan, needed by either
    // form of K4; redone below if lake removal or a renumbering changes the forest / the numbering first)
    const bool overlap = c->opt_overlap != 0 && c->opt_profile < 2;
    bool keys_ready = false;
    FL_RC(read_flags_begin(c));
    if (overlap && !c->need_rebuild) {
        LAUNCH_N(k_seg_keys, n, n, L_(c).recv, c->d_depth);
        FL_CK(fl_inclusive_max(c->d_tmp, c->tmp_bytes, c->d_depth, c->d_sg_head, n, c->stream, false));
        keys_ready = true;
    }
    FL_RC(read_flags_end(c));
    const bool has_lake = c->h_flags[FL_FLAG_LAKE] != 0;
    const uint32_t n_chg = c->h_flags[FL_FLAG_NCHG];
    FL_RC(stage_mark(c, 2));
    c->stages_valid = false;
    if (has_lake) {
        Layout& L = L_(c);
        FL_RC(run_labels(c));
        FL_RC(run_lakes(c));
        FL_CK(fl_memset(L.cmask, 0, sizeof(uint32_t) * n, c->stream));
        LAUNCH_N(k_childmask, n, n, L.row_ptr, L.col, L.rev, L.recv, L.cmask);
    }
    FL_RC(stage_mark(c, 3));

    const bool periodic = c->opt_rebuild_every > 0 && (it % (uint32_t)c->opt_rebuild_every) == 0;
    const bool rebuilt = c->need_rebuild || periodic || c->opt_rebuild_every == 1;
    if (rebuilt) {
        if (it == 0) {  // no previous areas yet: rank the children by subtree size (fl_flow.cuh)
            Layout& L0 = L_(c);
            LAUNCH_N(k_subtree_init, n, n, L0.cmask, c->d_state, c->d_nwait);
            LAUNCH_N(k_subtree_sweep, n, n, L0.recv, L0.cmask, c->d_state, c->d_nwait);
            LAUNCH_N(k_subtree_weight, n, n, c->d_state, c->d_A);
            c->stats.n_order += 3;
        }
        FL_RC(k_begin(c));
        FL_RC(rebuild_layout_flow(c, c->d_A));
        FL_RC(k_end(c, FASTLEM_K_REBUILD));
    }
    c->need_rebuild = false;
    if (has_lake || rebuilt) keys_ready = false;  // the forest / the numbering changed after the keys were taken
    FL_RC(stage_mark(c, 7));  // end of the layout rebuild
    Layout& L = L_(c);
    const bool incr = track && !has_lake && !rebuilt &&
                      (unsigned long long)n_chg * (unsigned long long)c->opt_incr_div <= (unsigned long long)n;

    // K4 (fl_flow.cuh): segment bookkeeping, then the two dataflow passes
    FlFlow f;
    f.n = n; f.row_ptr = L.row_ptr; f.col = L.col; f.recv = L.recv; f.cmask = L.cmask; f.areas = L.areas;
    f.A = c->d_A; f.nwait = c->d_nwait; f.seg_head = c->d_sg_head; f.seg_tail = c->d_sg_tail;
    f.seg_wait = c->d_sg_wait; f.seg_done = c->d_sg_done; f.state = c->d_state; f.pre = c->d_pre;
    f.post1 = c->d_post1; f.post2 = c->d_post2; f.xpost = c->d_xpost; f.hpre = c->d_hpre; f.xbuf = c->d_xbuf; f.hbuf = c->d_hbuf;
    f.hgt = c->d_hgt; f.flags = c->d_flags; f.parked = c->d_parked; f.counters = c->d_flags + FL_FLAG_PARKED;
    f.park_after = (uint32_t)c->opt_park_after; f.stats = c->d_flow_stats; f.tlog = c->d_tlog;
    f.hsuf = c->d_hsuf; f.dirty_from = nullptr; f.rlist = c->d_rlist; f.slist = c->d_slist;
    if (!incr) {
        FL_CK(fl_memset(c->d_state, 0, sizeof(uint32_t) * n, c->stream));
        FL_CK(fl_memset(c->d_nwait, 0, sizeof(uint32_t) * n, c->stream));
        FL_CK(fl_memset(c->d_sg_wait, 0, sizeof(uint32_t) * n, c->stream));
        FL_CK(fl_memset(c->d_sg_done, 0, sizeof(uint32_t) * n, c->stream));
        LAUNCH_N(k_count_waits, n, n, L.recv, L.cmask, L.areas, c->d_nwait, c->d_A, c->d_hgt, c->d_hsuf);
        if (!keys_ready) {
            LAUNCH_N(k_seg_keys, n, n, L.recv, c->d_depth);
            FL_CK(fl_inclusive_max(c->d_tmp, c->tmp_bytes, c->d_depth, c->d_sg_head, n, c->stream, false));
        }
        LAUNCH_N(k_seg_prepare, n, f, c->d_sg_tail, c->d_sg_wait);
        FL_RC(k_begin(c));
        LAUNCH_N(k_area_flow, n, f);
        FL_RC(k_end(c, FASTLEM_K_AREA_FLOW));
        if (f.park_after) {  // the parked (long) climbs, one warp each (persistent grid)
            FL_RC(k_begin(c));
            FL_LAUNCH(k_area_flow_long, (unsigned)c->sm_count * 4u, 256, c->stream, f);
            FL_RC(k_end(c, FASTLEM_K_AREA_FLOW_LONG));
            c->stats.kernel_launches++;
        }
        c->stats.n_area += 6;
        c->k4_valid = true;
        c->k4_last_full = true;
    } else {
        if (c->k4_last_full) {  // the full pass leaves its counters behind
            FL_CK(fl_memset(c->d_nwait, 0, sizeof(uint32_t) * n, c->stream));
            FL_CK(fl_memset(c->d_sg_wait, 0, sizeof(uint32_t) * n, c->stream));
            FL_CK(fl_memset(c->d_sg_done, 0, sizeof(uint32_t) * n, c->stream));
            FL_CK(fl_memset(c->d_dirty_from, 0, sizeof(uint32_t) * n, c->stream));
            c->k4_last_full = false;
        }
        if (n_chg) {
            f.dirty_from = c->d_dirty_from;
            if (!keys_ready) {
                LAUNCH_N(k_seg_keys, n, n, L.recv, c->d_depth);
                FL_CK(fl_inclusive_max(c->d_tmp, c->tmp_bytes, c->d_depth, c->d_sg_head, n, c->stream, false));
            }
            FL_LAUNCH(k_incr_mark, blocks_for(n_chg, 128), 128, c->stream, f, n_chg, c->d_chg_node, c->d_chg_old);
            const unsigned wide = (unsigned)c->sm_count * 8u;
            FL_LAUNCH(k_incr_prepare, wide, 128, c->stream, f);
            FL_RC(k_begin(c));
            FL_LAUNCH(k_incr_start, wide * 2u, 64, c->stream, f);
            FL_RC(k_end(c, FASTLEM_K_INCR_START));
            if (f.park_after) {
                FL_RC(k_begin(c));
                FL_LAUNCH(k_area_flow_long, (unsigned)c->sm_count * 4u, 256, c->stream, f);
                FL_RC(k_end(c, FASTLEM_K_AREA_FLOW_LONG));
            }
            FL_LAUNCH(k_incr_cleanup, wide, 256, c->stream, f);
            const uint64_t nl = f.park_after ? 7 : 6;
            c->stats.kernel_launches += nl;
            c->stats.n_area += nl;
        }
        c->stats.incremental_iterations++;
    }
    FL_RC(stage_mark(c, 8));  // end of K4

    uint32_t launched = 0;
    if (c->opt_k5_split) {
        // K5 split by nesting height (fl_elev.cuh): nothing is sorted, nothing is read back before the sweep
        FL_RC(stage_mark(c, 5));
        uint32_t cut = c->opt_k5_cut >= 0 ? (uint32_t)c->opt_k5_cut : c->k5_cut;
        if (c->opt_k5_cut < 0 && !c->k5_cut_known) cut = n <= (uint32_t)c->opt_k5_top_cap ? 0u : 4u;
        if (cut > FL_CUT_MAX) cut = FL_CUT_MAX;
        FlSplit e;
        e.n = n; e.row_ptr = L.row_ptr; e.col = L.col; e.rev = L.rev; e.recv = L.recv; e.cmask = L.cmask;
        e.is_outlet = L.is_outlet; e.hgt = c->d_hgt; e.drecv = L.drecv; e.erod = L.erod; e.A = c->d_A;
        e.uplift = L.uplift; e.tan_slope = c->has_tan ? L.tan : nullptr; e.tcel = c->d_tcel; e.elev = L.elev; e.rt = c->d_rt;
        e.root_of = c->d_root_of; e.pmask = c->d_pmask; e.queue = c->d_queue; e.epoch = ++c->push_epoch; e.cut = cut;
        e.flags = c->d_flags; e.low_list = c->d_low_list;
        FL_RC(k_begin(c));
#ifdef FL_EMU
        LAUNCH_N(k_elev_plan, n, e);
#else
        FL_LAUNCH(k_elev_plan, (n + FL_LOW_CHUNK - 1u) / FL_LOW_CHUNK, 256, c->stream, e);
        c->stats.kernel_launches++;
#endif
        FL_RC(k_end(c, FASTLEM_K_ELEV_PLAN));
        FL_RC(k_begin(c));
#ifdef FL_EMU
        FL_LAUNCH(k_elev_top, 1, 1, c->stream, e);
#else
        {
            int64_t blocks = c->opt_k5_top_blocks > 0 ? c->opt_k5_top_blocks : (int64_t)c->sm_count;
            if (blocks > (int64_t)c->top_blocks) blocks = c->top_blocks;
            FL_LAUNCH(k_elev_top, (unsigned)blocks, 256, c->stream, e);
        }
#endif
        FL_RC(k_end(c, FASTLEM_K_ELEV_TOP));
        c->stats.kernel_launches++;
        for (uint32_t lv = cut; lv-- > 0;) {
            FL_RC(k_begin(c));
#ifdef FL_EMU
            LAUNCH_N(k_elev_low, n, e, lv);
#else
            {   // persistent grid over the height's list; never more blocks than the list can have groups of 256 heads
                const size_t most = ((size_t)n / (lv + 1u) + 256u) / 256u;
                const unsigned blocks = (unsigned)(most < (size_t)c->low_blocks ? most : (size_t)c->low_blocks);
                FL_LAUNCH(k_elev_low, blocks, 256, c->stream, e, lv, fl_low_region(n, lv));
            }
            c->stats.kernel_launches++;
#endif
            FL_RC(k_end(c, FASTLEM_K_ELEV_LOW));
        }
        c->stats.n_elevation += 2 + cut;
#ifdef FL_EMU
        // emulation only: every push mask must have been consumed (a mask left behind = a segment at or above the cut
        // whose receiver's segment lies below it, i.e. the nesting heights are not strictly decreasing downwards)
        for (uint32_t i = 0; i < n; ++i)
            if (c->d_pmask[i]) return fail(c, FASTLEM_E_STATE, "K5: a push mask was not consumed (nesting heights inconsistent)");
#endif
    } else {
        // order the segment heads by descending nesting height (exact for the current forest; an upper bound of the
        // largest height after incremental passes -- heights that no longer occur are empty levels)
        // Keys are base - height with a fixed base, so the sort does not have to wait for the host to learn the largest
        // height: one read-back after the level offsets brings everything.  Only when segments nest deeper than the fixed
        // base (stale numbering in the first iterations) the ordering is redone with the exact height as base.
        uint32_t key_base = (uint32_t)c->opt_key_base;
        uint32_t* hbuf = c->h_offs_k;
        uint32_t maxh = 0;
        for (int attempt = 0;; ++attempt) {
            int bits = 8;
            if (attempt || key_base != FL_KEY_BASE) {
                bits = 1;
                while (bits < 32 && (1ull << bits) <= (unsigned long long)key_base + 1ull) ++bits;
            }
            if (attempt) {
                FL_CK(fl_memset(c->d_flags + FL_FLAG_BROKEN, 0, sizeof(uint32_t), c->stream));
                FL_CK(fl_d2d(c->d_flags + FL_FLAG_MAXDEPTH, c->d_flags + FL_FLAG_K4MAXH, sizeof(uint32_t), c->stream));
            }
            LAUNCH_N(k_flow_sort_keys, n, n, c->d_hgt, key_base, c->d_depth, c->d_flags);
            FL_CK(fl_sort_pairs(c->d_tmp, c->tmp_bytes, c->d_depth, c->d_sorted, c->d_iota, c->d_order, n, bits, c->stream,
                                false));
            FL_CK(fl_memset(c->d_flags + FL_FLAG_MAXDEPTH, 0xFF, sizeof(uint32_t), c->stream));
            FL_CK(fl_memset(c->d_offs, 0xFF, sizeof(uint32_t) * ((size_t)key_base + 2), c->stream));
            LAUNCH_N(k_level_offsets, n, n, c->d_sorted, c->d_offs, c->d_flags);
            FL_CK(fl_d2h(hbuf, c->d_offs, sizeof(uint32_t) * ((size_t)key_base + 2), c->stream));
            FL_RC(read_flags(c));
            if (c->h_flags[FL_FLAG_BROKEN] & 1u)
                return fail(c, FASTLEM_E_STATE, "K4: a climb met an unpublished site (internal error)");
            maxh = c->h_flags[FL_FLAG_K4MAXH];
            if (incr && c->prev_maxh > maxh) maxh = c->prev_maxh;
            // redo with the exact base when a key overflowed the fixed base, or when the bound carried over from the
            // previous iteration (incremental passes only see the heights of the dirty segments) lies above it
            if (!(c->h_flags[FL_FLAG_BROKEN] & 4u) && maxh <= key_base) break;
            if (attempt) return fail(c, FASTLEM_E_STATE, "flow: height bookkeeping broke (keys)");
            key_base = maxh;  // deeper than the fixed base: exact base, more key bits, the large host buffer
            FL_RC(ensure_h_offs(c));
            hbuf = c->h_offs;
        }
        uint32_t* const hoffs = hbuf + (key_base - maxh);  // hoffs[g]: first head of height maxh - g
        const uint32_t last_abs = c->h_flags[FL_FLAG_MAXDEPTH];  // = key_base - (smallest height that occurs)
        if (last_abs == FL_NONE || last_abs < key_base - maxh || last_abs > key_base || (!incr && last_abs != key_base))
            return fail(c, FASTLEM_E_STATE, "flow: height bookkeeping broke");
        const uint32_t last_key = last_abs - (key_base - maxh);
        const uint32_t n_heads = c->h_flags[FL_FLAG_REACHED];
        for (uint32_t g = last_key + 2; g <= maxh + 1; ++g) hoffs[g] = n_heads;
        for (uint32_t g = maxh + 1; g-- > 0;)
            if (hoffs[g] == FL_NONE) hoffs[g] = hoffs[g + 1];
        if (hoffs[0] == FL_NONE) hoffs[0] = 0;
        c->stats.n_order += 3;
        c->stats.path_levels = maxh + 1;
        c->stats.paths = n_heads;
        c->prev_maxh = maxh;
        if (rebuilt) { c->segs_at_rebuild = n_heads; c->maxh_at_rebuild = maxh; }
        else if (c->opt_rebuild_every == 0 &&
                 ((unsigned long long)n_heads * 100ull >
                      (unsigned long long)c->segs_at_rebuild * (100ull + rebuild_growth_of(c)) ||
                  (unsigned long long)maxh * 100ull >
                      (unsigned long long)c->maxh_at_rebuild * rebuild_height_of(c) + 200ull))
            c->need_rebuild = true;  // the numbering has degraded: renumber in the next iteration
        if (c->trace_iters && it < 400u)
            std::fprintf(stderr, "[fastlem trace] it %u: chg %u incr %d rebuilt %d heads %u maxh %u next_rebuild %d\n", it, n_chg,
                         (int)incr, (int)rebuilt, n_heads, maxh, (int)c->need_rebuild);
        FL_RC(stage_mark(c, 5));  // end of the head ordering

        // K5: one launch per nesting height, outermost segments first
        FlElev e;
        LAUNCH_N(k_celerity_term, n, n, L.erod, c->d_A, L.drecv, c->d_tcel);
        e.n = n; e.recv = L.recv; e.drecv = L.drecv; e.tcel = c->d_tcel; e.uplift = L.uplift;
        e.tan_slope = c->has_tan ? L.tan : nullptr; e.is_outlet = L.is_outlet; e.elev = L.elev; e.rt = c->d_rt;
        e.root_of = c->d_root_of; e.flags = c->d_flags; e.lvl = L.lvl; e.lvl_value = 0;
        // the sparse levels at the top of the forest go out as ONE launch (ticket order = level order)
        uint32_t g_first = 0;
        if (c->opt_fuse_levels) {
            uint32_t gf = 0;
            while (gf <= maxh && hoffs[gf + 1] - hoffs[gf] <= FL_WARP_LEVEL_MAX) ++gf;
            const uint32_t total = hoffs[gf];
            if (gf >= 2 && total > 0) {
                FlFused u;
                u.count = total; u.heads = c->d_order; u.seg_head = c->d_sg_head; u.ticket_of = c->d_ticket_of;
                u.done = c->d_fdone; u.next_ticket = c->d_flags + FL_FLAG_TICKET; u.lvl_of = c->d_flvl;
                LAUNCH_N(k_fused_index, total, total, c->d_order, c->d_hgt, c->d_ticket_of, c->d_flvl, c->d_fdone);
    #ifdef FL_EMU
                const unsigned blocks = blocks_for(total, 128);  // emulation: one thread per ticket, in ticket order
    #else
                unsigned blocks = blocks_for(total * 32u, 128);
                const unsigned cap = (unsigned)c->sm_count * 8u;
                if (blocks > cap) blocks = cap;
    #endif
                FL_LAUNCH(k_elev_flow_fused, blocks, 128, c->stream, u, e);
                launched += 1;
                g_first = gf;
            }
        }
        for (uint32_t g = g_first; g <= maxh; ++g) {
            const uint32_t b = hoffs[g], cnt = hoffs[g + 1] - b;
            if (!cnt) continue;
            ++launched;
            e.lvl_value = maxh - g;
            if (cnt <= FL_WARP_LEVEL_MAX)  // few segments: a warp each
                FL_LAUNCH(k_elev_flow_warps, blocks_for(cnt * 32u, 128), 128, c->stream, b, cnt, c->d_order, e);
            else
                FL_LAUNCH(k_elev_flow, blocks_for(cnt, 128), 128, c->stream, b, cnt, c->d_order, e);
        }
    }
    c->stats.kernel_launches += launched; c->stats.n_elevation += launched;
    FL_RC(stage_mark(c, 6));
    FL_RC(read_flags_begin(c));
    if (overlap && may_continue && c->opt_k5_split) {
        // While the flags travel to the host, the NEXT iteration's K1 is already running (the elevations it reads are
        // final).  It writes the alternate receiver buffers, so if this iteration turns out to be the last one
        // (`changed` clear) nothing of the run's state has been touched; the next iteration swaps the buffers in.
        const bool track_next = c->opt_incremental != 0 && c->opt_rebuild_every != 1;  // (k4_valid holds from here on)
        FL_CK(fl_memset(c->d_flags, 0, sizeof(uint32_t) * FL_N_FLAGS, c->stream));
        FL_CK(fl_memset(c->d_cmask_alt, 0, sizeof(uint32_t) * n, c->stream));
        if (c->opt_profile) FL_CK(fl_event_record(c->ev[ST_COUNT + 3], c->stream));
        FL_RC(launch_receivers(c, track_next, L.recv, c->d_recv_alt, c->d_drecv_alt, c->d_cmask_alt));
        if (c->opt_profile) FL_CK(fl_event_record(c->ev[ST_COUNT + 4], c->stream));
        c->k1_pre = true;
        c->k1_pre_track = track_next;
    }
    FL_RC(read_flags_end(c));
    FL_CK(fl_last_error());
    if (c->opt_k5_split) {
        if (c->h_flags[FL_FLAG_BROKEN] & 1u)
            return fail(c, FASTLEM_E_STATE, "K4: a climb met an unpublished site (internal error)");
        if (c->h_flags[FL_FLAG_BROKEN])
            return fail(c, FASTLEM_E_STATE,
                        "K5: the run queue broke (internal error; flags " + std::to_string(c->h_flags[FL_FLAG_BROKEN]) +
                            " tail " + std::to_string(c->h_flags[FLQ_TAIL]) + " head " + std::to_string(c->h_flags[FLQ_HEAD]) +
                            " done " + std::to_string(c->h_flags[FLQ_DONE]) + " iteration " + std::to_string(it) + ")");
        // bookkeeping of the numbering: nesting height met by K4 (at the roots it finished), segments of the forest
        uint32_t maxh = c->h_flags[FL_FLAG_MAXDEPTH];
        if (incr && c->prev_maxh > maxh) maxh = c->prev_maxh;
        const uint32_t* hist = c->h_flags + FLQ_HIST;
        uint32_t n_heads = 0;
        for (int b = 0; b < 32; ++b) n_heads += hist[b];
        // cut height of the next sweep: the smallest one that leaves at most top_cap segments to the queue
        {
            uint32_t above = 0, cut = FL_CUT_MAX;
            for (int b = 31; b >= (int)FL_CUT_MAX; --b) above += hist[b];
            while (cut > 0u && above + hist[cut - 1u] <= (uint32_t)c->opt_k5_top_cap) { above += hist[cut - 1u]; --cut; }
            c->k5_cut = cut;
            c->k5_cut_known = true;
        }
        c->stats.path_levels = maxh + 1;
        c->stats.paths = n_heads;
        c->prev_maxh = maxh;
        if (rebuilt) { c->segs_at_rebuild = n_heads; c->maxh_at_rebuild = maxh; }
        else if (c->opt_rebuild_every == 0 &&
                 ((unsigned long long)n_heads * 100ull >
                      (unsigned long long)c->segs_at_rebuild * (100ull + rebuild_growth_of(c)) ||
                  (unsigned long long)maxh * 100ull >
                      (unsigned long long)c->maxh_at_rebuild * rebuild_height_of(c) + 200ull))
            c->need_rebuild = true;
        if (c->trace_iters && it < 400u)
            std::fprintf(stderr, "[fastlem trace] it %u: chg %u incr %d rebuilt %d heads %u maxh %u cut %u top %u next_rebuild %d\n",
                         it, n_chg, (int)incr, (int)rebuilt, n_heads, maxh, c->k5_cut, c->h_flags[FLQ_TAIL],
                         (int)c->need_rebuild);
    } else if (c->h_flags[FL_FLAG_BROKEN])
        return fail(c, FASTLEM_E_STATE, "K5: a segment waited for its receiver's segment too long (internal error)");
    *changed_out = c->h_flags[FL_FLAG_CHANGED] != 0;
    if (c->opt_profile) {  // receivers | flags | lakes | rebuild + ordering | K4 | K5
        const int from[7] = {0, 1, 2, 3, 8, 7, 5}, to[7] = {1, 2, 3, 7, 5, 8, 6};
        double* acc[7] = {&c->stats.ms_receivers, &c->stats.ms_labels, &c->stats.ms_lakes, &c->stats.ms_order,
                          &c->stats.ms_order,     &c->stats.ms_area,   &c->stats.ms_elevation};
        for (int k = 0; k < 7; ++k) {
            float ms = 0.f;
            FL_CK(fl_event_elapsed(&ms, c->ev[from[k]], c->ev[to[k]]));
            *acc[k] += ms;
        }
    }
    return FASTLEM_OK;
}

// start of a run: working numbering = the caller's
int reset_layout(fastlem_ctx* c) {
    const uint32_t n = c->n, nnz = c->nnz;
    c->cur = 0;
    Layout& L = c->lay[0];
    const Layout& O = c->orig;
    FL_CK(fl_d2d(L.row_ptr, O.row_ptr, sizeof(uint32_t) * ((size_t)n + 1), c->stream));
    if (nnz) {
        FL_CK(fl_d2d(L.col, O.col, sizeof(uint32_t) * nnz, c->stream));
        FL_CK(fl_d2d(L.dist, O.dist, sizeof(double) * nnz, c->stream));
        FL_CK(fl_d2d(L.rev, O.rev, nnz, c->stream));
    }
    FL_CK(fl_d2d(L.areas, O.areas, sizeof(double) * n, c->stream));
    FL_CK(fl_d2d(L.erod, O.erod, sizeof(double) * n, c->stream));
    FL_CK(fl_d2d(L.uplift, O.uplift, sizeof(double) * n, c->stream));
    if (c->has_tan) FL_CK(fl_d2d(L.tan, O.tan, sizeof(double) * n, c->stream));
    FL_CK(fl_d2d(L.is_outlet, O.is_outlet, n, c->stream));
    FL_CK(fl_d2d(L.elev, c->d_init, sizeof(double) * n, c->stream));
    LAUNCH_N(k_iota, n, n, L.orig_of);
    FL_CK(fl_memset(L.lvl, 0, sizeof(uint32_t) * n, c->stream));
    FL_CK(fl_memset(c->d_pmask, 0, sizeof(uint32_t) * n, c->stream));
    c->k5_cut_known = false;
    c->prev_maxh = 0;
    c->k4_valid = false;
    c->k4_last_full = true;
    if (c->rank_ready) {
        FL_CK(fl_d2d(L.rank, O.rank, sizeof(uint32_t) * n, c->stream));
        FL_CK(fl_memset(c->d_rank_to_node, 0xFF, sizeof(uint32_t) * n, c->stream));
        LAUNCH_N(k_rank_inverse, n, n, L.rank, c->d_rank_to_node);
    }
    c->layout_valid = true;
    return FASTLEM_OK;
}

// every device buffer of a graph (see fastlem_ctx::slab)
int alloc_graph_buffers(fastlem_ctx* c, uint32_t n, uint32_t nnz) {
    const size_t n1 = (size_t)n + 1;
    Layout* sets[3] = {&c->orig, &c->lay[0], &c->lay[1]};
    for (Layout* S : sets) {
        FL_CK(dalloc(c, S->row_ptr, n1));
        FL_CK(dalloc(c, S->col, nnz));
        FL_CK(dalloc(c, S->dist, nnz));
        FL_CK(dalloc(c, S->rev, nnz));
        FL_CK(dalloc(c, S->areas, n));
        FL_CK(dalloc(c, S->erod, n));
        FL_CK(dalloc(c, S->uplift, n));
        FL_CK(dalloc(c, S->is_outlet, n));
        FL_CK(dalloc(c, S->rank, n));
    }
    for (int k = 0; k < 2; ++k) {
        Layout& S = c->lay[k];
        FL_CK(dalloc(c, S.orig_of, n));
        FL_CK(dalloc(c, S.elev, n));
        FL_CK(dalloc(c, S.drecv, n));
        FL_CK(dalloc(c, S.recv, n));
        FL_CK(dalloc(c, S.cmask, n));
        FL_CK(dalloc(c, S.lvl, n));
    }
    FL_CK(dalloc(c, c->d_recv_alt, n));
    FL_CK(dalloc(c, c->d_cmask_alt, n));
    FL_CK(dalloc(c, c->d_drecv_alt, n));
    FL_CK(dalloc(c, c->d_init, n));
    FL_CK(dalloc(c, c->d_rank_to_node, n));
    FL_CK(dalloc(c, c->d_pd, n));
    FL_CK(dalloc(c, c->d_pd2, n));
    FL_CK(dalloc(c, c->d_lake_key, n));
    FL_CK(dalloc(c, c->d_label, n));
    FL_CK(dalloc(c, c->d_depth, n));
    FL_CK(dalloc(c, c->d_ids, n));
    FL_CK(dalloc(c, c->d_sorted, n));
    FL_CK(dalloc(c, c->d_order, n));
    FL_CK(dalloc(c, c->d_offs, (size_t)n + 2 > FL_KEY_BASE + 2 ? (size_t)n + 2 : (size_t)FL_KEY_BASE + 2));
    FL_CK(dalloc(c, c->d_A, n));
    FL_CK(dalloc(c, c->d_rt, n));
    FL_CK(dalloc(c, c->d_root_of, n));
    FL_CK(dalloc(c, c->d_heavy, n));
    FL_CK(dalloc(c, c->d_plen, n));
    FL_CK(dalloc(c, c->d_len_sorted, n));
    FL_CK(dalloc(c, c->d_hrank, n));
    FL_CK(dalloc(c, c->d_seg_head, n1));
    FL_CK(dalloc(c, c->d_newpos, n));
    FL_CK(dalloc(c, c->d_deg_new, n1));
    FL_CK(dalloc(c, c->d_state, n));
    FL_CK(dalloc(c, c->d_pre, n));
    FL_CK(dalloc(c, c->d_post1, n));
    FL_CK(dalloc(c, c->d_post2, n));
    FL_CK(dalloc(c, c->d_xbuf, n));
    FL_CK(dalloc(c, c->d_xpost, (size_t)n * FL_XPOST));
    FL_CK(dalloc(c, c->d_hbuf, n));
    FL_CK(dalloc(c, c->d_hgt, n));
    FL_CK(dalloc(c, c->d_hpre, n));
    FL_CK(dalloc(c, c->d_iota, n));
    FL_CK(dalloc(c, c->d_parked, (size_t)n / 4 + 64));
    FL_CK(dalloc(c, c->d_nwait, n));
    FL_CK(dalloc(c, c->d_sg_head, n));
    FL_CK(dalloc(c, c->d_sg_tail, n));
    FL_CK(dalloc(c, c->d_sg_wait, n));
    FL_CK(dalloc(c, c->d_sg_done, n));
    FL_CK(dalloc(c, c->d_low_list, fl_low_region(n, FL_CUT_MAX)));
    FL_CK(dalloc(c, c->d_queue, n));
    FL_CK(dalloc(c, c->d_pmask, n));
    FL_CK(dalloc(c, c->d_ticket_of, n));
    FL_CK(dalloc(c, c->d_fdone, n));
    FL_CK(dalloc(c, c->d_flvl, n));
    FL_CK(dalloc(c, c->d_hsuf, n));
    FL_CK(dalloc(c, c->d_dirty_from, n));
    FL_CK(dalloc(c, c->d_rlist, n));
    FL_CK(dalloc(c, c->d_slist, n));
    FL_CK(dalloc(c, c->d_chg_node, n));
    FL_CK(dalloc(c, c->d_chg_old, n));
#ifdef FL_FLOW_STATS
    FL_CK(dalloc(c, c->d_tlog, (size_t)n * 4));
#endif
    FL_CK(dalloc(c, c->d_flow_stats, 32));
    FL_CK(dalloc(c, c->d_tcel, n));
    FL_CK(dalloc(c, c->d_out_f64, n));
    FL_CK(dalloc(c, c->d_out_u32, n));
    FL_CK(dalloc(c, c->d_recv0, n));
    FL_CK(dalloc(c, c->d_label0, n));
    return FASTLEM_OK;
}

}  // namespace

extern "C" {

int fastlem_trim_memory(int device_ordinal) {
    if (fl_set_device(device_ordinal) != cudaSuccess) return FASTLEM_E_CUDA;
    return fl_trim_pool() == cudaSuccess ? FASTLEM_OK : FASTLEM_E_CUDA;
}

const char* fastlem_version(void) {
#ifdef FL_EMU
    return "fastlem_b200 0.2.0 emu";
#else
    return "fastlem_b200 0.2.0 sm_100a";
#endif
}

int fastlem_create(fastlem_ctx** out, int device_ordinal) {
    if (!out) return FASTLEM_E_INVALID;
    *out = nullptr;
    fastlem_ctx* c = new (std::nothrow) fastlem_ctx();
    if (!c) return FASTLEM_E_NOMEM;
    c->device = device_ordinal;
    cudaError_t e = fl_set_device(device_ordinal);
    if (e == cudaSuccess) e = fl_stream_create(&c->stream);
    if (e != cudaSuccess) {
        // no CPU fallback: without a usable CUDA device there is no ctx
        std::fprintf(stderr, "fastlem_b200: cannot create context on CUDA device %d: %s\n", device_ordinal,
                     cudaGetErrorString(e));
        delete c;
        return FASTLEM_E_CUDA;
    }
    c->stream_ok = true;
    // one pinned block: flag words | level offsets of sweep 3 | round counts
    void* hf = nullptr;
    if (fl_malloc_host(&hf, sizeof(uint32_t) * (FL_N_FLAGS + (FL_KEY_BASE + 2) + (FL_MAX_ROUNDS + 2))) != cudaSuccess) {
        fl_stream_destroy(c->stream);
        delete c;
        return FASTLEM_E_NOMEM;
    }
    c->h_flags = (uint32_t*)hf;
    c->h_offs_k = c->h_flags + FL_N_FLAGS;
    c->h_rounds = c->h_offs_k + (FL_KEY_BASE + 2);
    bool ok = true;
    for (int k = 0; k < ST_COUNT + 6; ++k) ok = ok && fl_event_create(&c->ev[k]) == cudaSuccess;
    for (int k = 0; k < 2; ++k) ok = ok && fl_event_create(&c->ev_run[k]) == cudaSuccess;
    for (int k = 0; k < 2; ++k) ok = ok && fl_event_create(&c->ev_k[k]) == cudaSuccess;
    void* fl = nullptr;
    ok = ok && fl_malloc(&fl, sizeof(uint32_t) * FL_N_FLAGS) == cudaSuccess;
    c->d_flags = (uint32_t*)fl;
    if (!ok) {
        fastlem_destroy(c);
        return FASTLEM_E_CUDA;
    }
    *out = c;
    return FASTLEM_OK;
}

void fastlem_destroy(fastlem_ctx* c) {
    if (!c) return;
    fl_set_device(c->device);
    FlTrace tr("destroy");
    free_all(c);
    tr.mark("free_all");
    if (c->d_tmp) fl_free(c->d_tmp, c->stream);
    if (c->d_flags) fl_free(c->d_flags, c->stream);
    if (c->h_flags) fl_free_host(c->h_flags);

    for (int k = 0; k < ST_COUNT + 6; ++k)
        if (c->ev[k]) fl_event_destroy(c->ev[k]);
    for (int k = 0; k < 2; ++k)
        if (c->ev_run[k]) fl_event_destroy(c->ev_run[k]);
    for (int k = 0; k < 2; ++k)
        if (c->ev_k[k]) fl_event_destroy(c->ev_k[k]);
    if (c->stream_ok) fl_stream_destroy(c->stream);
    delete c;
}

const char* fastlem_last_error(const fastlem_ctx* c) { return c ? c->err.c_str() : "null context"; }

int fastlem_get_device(const fastlem_ctx* c) { return c ? c->device : -1; }

int fastlem_set_option(fastlem_ctx* c, const char* name, int64_t value) {
    if (!c || !name) return FASTLEM_E_INVALID;
    std::string s(name);
    if (s == "profile") c->opt_profile = value < 0 ? 0 : (value > 2 ? 2 : (int)value);
    else if (s == "keep_stages") c->opt_keep = value != 0;
    else if (s == "sweep") {
        if (value < 0 || value > 3)
            return fail(c, FASTLEM_E_INVALID,
                        "option sweep: 0 (levels), 1 (paths, thread per path), 2 (paths, warp per long path), 3 (dataflow)");
        c->opt_sweep = value;
    } else if (s == "park_after") {
        if (value != 0 && value < 4) return fail(c, FASTLEM_E_INVALID, "option park_after: 0 (never) or >= 4");
        c->opt_park_after = value;
    } else if (s == "flood_device") {
        c->opt_flood_device = value != 0;
        c->rank_ready = false;
    } else if (s == "overlap") {
        c->opt_overlap = value != 0;
    } else if (s == "k1_bulk") {
        c->opt_k1_bulk = value != 0;
    } else if (s == "outlet_closed_form") {
        c->opt_outlet_closed_form = value != 0;
        c->rank_ready = false;
    } else if (s == "incremental") {
        c->opt_incremental = value != 0;
    } else if (s == "incr_div") {
        if (value < 1) return fail(c, FASTLEM_E_INVALID, "option incr_div: >= 1");
        c->opt_incr_div = value;
    } else if (s == "first_flow") {
        c->opt_first_flow = value != 0;
    } else if (s == "key_base") {
        if (value < 1 || value > (int64_t)FL_KEY_BASE) return fail(c, FASTLEM_E_INVALID, "option key_base: 1..254");
        c->opt_key_base = value;
    } else if (s == "k5_split") {
        c->opt_k5_split = value != 0;
    } else if (s == "k5_cut") {
        if (value < -1 || value > (int64_t)FL_CUT_MAX) return fail(c, FASTLEM_E_INVALID, "option k5_cut: -1 (auto) or 0..8");
        c->opt_k5_cut = value;
    } else if (s == "k5_top_cap") {
        if (value < 0) return fail(c, FASTLEM_E_INVALID, "option k5_top_cap: >= 0");
        c->opt_k5_top_cap = value;
    } else if (s == "k5_top_blocks") {
        if (value < 0) return fail(c, FASTLEM_E_INVALID, "option k5_top_blocks: 0 (one per SM) or a block count");
        c->opt_k5_top_blocks = value;
    } else if (s == "fuse_levels") {
        c->opt_fuse_levels = value != 0;
    } else if (s == "rebuild_height") {
        if (value != 0 && value < 100) return fail(c, FASTLEM_E_INVALID, "option rebuild_height: 0 (by model size) or percent >= 100");
        c->opt_rebuild_height = value;
    } else if (s == "rebuild_growth") {
        if (value < 0) return fail(c, FASTLEM_E_INVALID, "option rebuild_growth: 0 (by model size) or percent >= 1");
        c->opt_rebuild_growth = value;
    } else if (s == "rebuild_every") {
        if (value < 0) return fail(c, FASTLEM_E_INVALID, "option rebuild_every: 0 (adaptive) or a positive period");
        c->opt_rebuild_every = value;
    } else return fail(c, FASTLEM_E_INVALID, "unknown option: " + s);
    return FASTLEM_OK;
}

int fastlem_set_graph(fastlem_ctx* c, uint32_t n, const uint32_t* row_ptr, const uint32_t* col, const double* dist,
                      const double* areas) {
    if (!c) return FASTLEM_E_INVALID;
    if (!row_ptr || !areas) return fail(c, FASTLEM_E_INVALID, "set_graph: null pointer");
    if (n == 0 || n >= FL_NONE - 1) return fail(c, FASTLEM_E_INVALID, "set_graph: n must be in [1, 2^32-3]");
    if (row_ptr[0] != 0) return fail(c, FASTLEM_E_INVALID, "set_graph: row_ptr[0] must be 0");
    for (uint32_t i = 0; i < n; ++i)
        if (row_ptr[i + 1] < row_ptr[i]) return fail(c, FASTLEM_E_INVALID, "set_graph: row_ptr must be non-decreasing");
    const uint32_t nnz = row_ptr[n];
    if (nnz && (!col || !dist)) return fail(c, FASTLEM_E_INVALID, "set_graph: null col/dist");
    for (uint32_t s = 0; s < nnz; ++s)
        if (col[s] >= n) return fail(c, FASTLEM_E_INVALID, "set_graph: neighbour index out of range");
    FL_CK(fl_set_device(c->device));
    double t0 = wall_ms();
    FlTrace tr("set_graph");
    free_all(c);
    tr.mark("free previous");
    c->n = n;
    c->nnz = nnz;
    const size_t n1 = (size_t)n + 1;
    c->slab_dry = true;  // sizes first, then one allocation, then the same calls carve it up
    c->slab_need = 0;
    FL_RC(alloc_graph_buffers(c, n, nnz));
    c->slab_dry = false;
    {
        void* v = nullptr;
        FL_CK(fl_malloc(&v, c->slab_need));
        c->slab = (unsigned char*)v;
        c->slab_size = c->slab_need;
        c->slab_used = 0;
    }
    FL_RC(alloc_graph_buffers(c, n, nnz));
    tr.mark("layout allocations");
    FL_CK(fl_h2d(c->orig.row_ptr, row_ptr, sizeof(uint32_t) * n1, c->stream));
    if (nnz) {
        FL_CK(fl_h2d(c->orig.col, col, sizeof(uint32_t) * nnz, c->stream));
        FL_CK(fl_h2d(c->orig.dist, dist, sizeof(double) * nnz, c->stream));
    }
    FL_CK(fl_h2d(c->orig.areas, areas, sizeof(double) * n, c->stream));
    FL_CK(fl_memset(c->d_flags, 0, sizeof(uint32_t) * FL_N_FLAGS, c->stream));
    LAUNCH_N(k_rev_slots, n, n, c->orig.row_ptr, c->orig.col, c->orig.dist, c->orig.rev, c->d_flags);
    FL_CK(fl_d2h(c->h_flags, c->d_flags, sizeof(uint32_t), c->stream));
    FL_CK(fl_memset(c->d_queue, 0, sizeof(FlQEntry) * n, c->stream));
    FL_CK(fl_memset(c->d_rt, 0, sizeof(double) * n, c->stream));  // (read, never used, for sites of trees without outlet)
    // the climbs of K4 request a batch's / a window's partial sums before they know which of its sites have any: sites that
    // never publish are read (and ignored) too -- defined values keep compute-sanitizer's initcheck clean
    FL_CK(fl_memset(c->d_pre, 0, sizeof(double) * n, c->stream));
    FL_CK(fl_memset(c->d_post1, 0, sizeof(double) * n, c->stream));
    FL_CK(fl_memset(c->d_post2, 0, sizeof(double) * n, c->stream));
    FL_CK(fl_memset(c->d_xpost, 0, sizeof(double) * n * FL_XPOST, c->stream));
    FL_CK(fl_memset(c->d_xbuf, 0, sizeof(double) * n, c->stream));
    FL_CK(fl_memset(c->d_hbuf, 0, sizeof(uint32_t) * n, c->stream));
    FL_CK(fl_memset(c->d_hpre, 0, sizeof(uint32_t) * n, c->stream));
    FL_CK(fl_memset(c->d_hsuf, 0, sizeof(uint32_t) * n, c->stream));
    c->push_epoch = 0;
#ifndef FL_EMU
    {
        // k_elev_top waits on queue entries: every block must be resident
        int occ = 0;
        c->top_blocks = fl_sm_count();
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_elev_top, 256, 0) != cudaSuccess || occ < 1)
            return fail(c, FASTLEM_E_CUDA, "set_graph: k_elev_top does not fit on an SM");
        c->top_blocks = occ * fl_sm_count();  // the most that can be resident
    }
    {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_elev_low, 256, 0) != cudaSuccess || occ < 1) occ = 1;
        c->low_blocks = occ * fl_sm_count();
    }
    {
        int occ = 0;
        c->k1_bulk_blocks = 0;
        if (cudaFuncSetAttribute(k_receivers_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FL_K1B_SMEM) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_receivers_bulk, 256, FL_K1B_SMEM) == cudaSuccess && occ > 0)
            c->k1_bulk_blocks = occ * fl_sm_count();
    }
#endif
#ifdef FL_FLOW_STATS
    FL_CK(fl_memset(c->d_tlog, 0, sizeof(unsigned long long) * 4 * n, c->stream));
#endif
    FL_CK(fl_memset(c->d_flow_stats, 0, 32 * sizeof(unsigned long long), c->stream));
    c->sm_count = fl_sm_count();
    LAUNCH_N(k_iota, n, n, c->d_iota);
    c->max_degree = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t d = row_ptr[i + 1] - row_ptr[i];
        if (d > c->max_degree) c->max_degree = d;
    }
    // CUB temp storage: the larger of the sort and the scan
    size_t sort_bytes = 0, scan_bytes = 0;
    FL_CK(fl_sort_pairs(nullptr, sort_bytes, c->d_depth, c->d_sorted, c->d_ids, c->d_order, n, 32, c->stream, true));
    FL_CK(fl_exclusive_sum(nullptr, scan_bytes, c->d_deg_new, c->d_seg_head, n + 1, c->stream, true));
    size_t max_bytes = 0;
    FL_CK(fl_inclusive_max(nullptr, max_bytes, c->d_deg_new, c->d_sg_head, n, c->stream, true));
    if (max_bytes > scan_bytes) scan_bytes = max_bytes;
    if (c->d_tmp) { fl_free(c->d_tmp, c->stream); c->d_tmp = nullptr; }
    c->tmp_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
    FL_CK(fl_malloc(&c->d_tmp, c->tmp_bytes));
    tr.mark("work allocations");
    FL_CK(fl_stream_sync(c->stream));
    tr.mark("copies + sync");
    FL_CK(fl_last_error());
    if (c->h_flags[0]) {
        const uint32_t fw = c->h_flags[0];
        free_all(c);
        return fail(c, FASTLEM_E_INVALID,
                    std::string("set_graph: the graph must be simple and symmetric with equal lengths in both directions (") +
                        ((fw & 1u) ? "self loop; " : "") + ((fw & 2u) ? "parallel edges; " : "") +
                        ((fw & 4u) ? "edge without its reverse; " : "") + ((fw & 8u) ? "lengths differ between the two directions; " : "") +
                        "terrain-graph's add_edge on a triangulation never produces these)");
    }
    c->has_graph = true;
    c->stats = fastlem_stats{};
    c->stats.ms_upload = wall_ms() - t0;
    return FASTLEM_OK;
}

int fastlem_set_parameters(fastlem_ctx* c, const double* initial_elevation, const double* erodibility,
                           const double* uplift_rate, const double* tan_max_slope, const uint32_t* outlets,
                           uint32_t n_outlets) {
    if (!c) return FASTLEM_E_INVALID;
    if (!c->has_graph) return fail(c, FASTLEM_E_STATE, "set_parameters: call fastlem_set_graph first (ModelNotSet)");
    if (!initial_elevation || !erodibility || !uplift_rate || (n_outlets && !outlets))
        return fail(c, FASTLEM_E_INVALID, "set_parameters: null pointer");
    const uint32_t n = c->n;
    // stream_tree.rs:101-107 outlet table (static across iterations, so built once).  An outlet listed twice would be
    // traversed twice by the reference (generator.rs:149: its basin's areas and response times accumulate twice); the
    // crate's own outlet lists (ascending is_outlet indices / the hull cycle, generator.rs:120-132) never repeat a
    // site, so duplicates are rejected instead of silently changing the result.
    std::vector<uint8_t> table(n, 0);
    for (uint32_t k = 0; k < n_outlets; ++k) {
        if (outlets[k] >= n) return fail(c, FASTLEM_E_INVALID, "set_parameters: outlet index out of range");
        if (table[outlets[k]]) return fail(c, FASTLEM_E_INVALID, "set_parameters: an outlet is listed twice");
        table[outlets[k]] = 1;
    }
    FL_CK(fl_set_device(c->device));
    double t0 = wall_ms();
    FL_CK(fl_h2d(c->d_init, initial_elevation, sizeof(double) * n, c->stream));
    FL_CK(fl_h2d(c->orig.erod, erodibility, sizeof(double) * n, c->stream));
    FL_CK(fl_h2d(c->orig.uplift, uplift_rate, sizeof(double) * n, c->stream));
    c->has_tan = tan_max_slope != nullptr;
    if (c->has_tan) {
        if (!c->orig.tan) {
            FL_CK(dalloc(c, c->orig.tan, n));
            FL_CK(dalloc(c, c->lay[0].tan, n));
            FL_CK(dalloc(c, c->lay[1].tan, n));
        }
        FL_CK(fl_h2d(c->orig.tan, tan_max_slope, sizeof(double) * n, c->stream));
    }
    FL_CK(fl_h2d(c->orig.is_outlet, table.data(), n, c->stream));
    FL_CK(fl_stream_sync(c->stream));
    // the flood order of lake removal is a function of (graph, outlets) only: keep it when an ensemble member changes
    // nothing but the per-site parameters
    const bool same_outlets = c->outlets.size() == n_outlets && std::equal(c->outlets.begin(), c->outlets.end(), outlets);
    c->outlets.assign(outlets, outlets + n_outlets);
    c->rank_ready = c->rank_ready && same_outlets;
    c->has_params = true;
    c->stats.ms_upload += wall_ms() - t0;
    return FASTLEM_OK;
}

int fastlem_run(fastlem_ctx* c, uint32_t max_iteration, uint32_t* iterations_done) {
    if (!c) return FASTLEM_E_INVALID;
    if (!c->has_graph) return fail(c, FASTLEM_E_STATE, "run: model not set (ModelNotSet)");
    if (!c->has_params) return fail(c, FASTLEM_E_STATE, "run: parameters not set (ParametersNotSet)");
    FL_CK(fl_set_device(c->device));
    // reset per-run stats, keep the one-off ones
    double up = c->stats.ms_upload, fr = c->stats.ms_flood_rank;
    const uint32_t fd = c->stats.flood_on_device, od = c->stats.outlet_ranks_on_device;
    c->stats = fastlem_stats{};
    c->stats.ms_upload = up;
    c->stats.ms_flood_rank = fr;
    c->stats.flood_on_device = fd;
    c->stats.outlet_ranks_on_device = od;
    FL_CK(fl_event_record(c->ev_run[0], c->stream));
    FL_RC(reset_layout(c));
    c->need_rebuild = true;
    c->k1_pre = false;
    uint32_t it = 0;
    while (it < max_iteration) {
        bool changed = false;
        // The first body runs level-synchronously: it yields the drainage areas that rank the heavy
        // children of the path layout from the second body on.
        const bool flow_ok = c->opt_sweep == 3 && c->max_degree <= 32;
        if (c->opt_sweep == 0 || (it == 0 && !(flow_ok && c->opt_first_flow))) FL_RC(iterate_levels(c, it == 0, &changed));
        else if (flow_ok) FL_RC(iterate_flow(c, it, it + 1u < max_iteration, &changed));
        else FL_RC(iterate_paths(c, &changed));
        ++it;
        if (!changed) break;
    }
    FL_CK(fl_event_record(c->ev_run[1], c->stream));
    FL_CK(fl_event_sync(c->ev_run[1]));
    float ms = 0.f;
    FL_CK(fl_event_elapsed(&ms, c->ev_run[0], c->ev_run[1]));
    c->stats.ms_run = ms;
    c->stats.iterations = it;
    if (iterations_done) *iterations_done = it;
    return FASTLEM_OK;
}

int fastlem_download_to_device(fastlem_ctx* c, double* device_out) {
    if (!c || !device_out) return FASTLEM_E_INVALID;
    if (!c->layout_valid) return fail(c, FASTLEM_E_STATE, "download: nothing to download (no run yet)");
    FL_CK(fl_set_device(c->device));
    LAUNCH_N(k_unpermute_f64, c->n, c->n, L_(c).orig_of, L_(c).elev, device_out);
    FL_CK(fl_stream_sync(c->stream));
    return FASTLEM_OK;
}

int fastlem_download(fastlem_ctx* c, double* elevations_out) {
    if (!c || !elevations_out) return FASTLEM_E_INVALID;
    if (!c->layout_valid) return fail(c, FASTLEM_E_STATE, "download: nothing to download (no run yet)");
    FL_CK(fl_set_device(c->device));
    double t0 = wall_ms();
    LAUNCH_N(k_unpermute_f64, c->n, c->n, L_(c).orig_of, L_(c).elev, c->d_out_f64);
    FL_CK(fl_d2h(elevations_out, c->d_out_f64, sizeof(double) * c->n, c->stream));
    FL_CK(fl_stream_sync(c->stream));
    c->stats.ms_download = wall_ms() - t0;
    return FASTLEM_OK;
}

int fastlem_generate(fastlem_ctx* c, uint32_t max_iteration, double* elevations_out, uint32_t* iterations_done) {
    if (!c || !elevations_out) return FASTLEM_E_INVALID;
    int rc = fastlem_run(c, max_iteration, iterations_done);
    if (rc) return rc;
    return fastlem_download(c, elevations_out);
}

int fastlem_get_stats(const fastlem_ctx* c, fastlem_stats* out) {
    if (!c || !out) return FASTLEM_E_INVALID;
    *out = c->stats;
    return FASTLEM_OK;
}

int fastlem_debug_fetch(fastlem_ctx* c, int stage, void* out, size_t bytes) {
    if (!c || !out) return FASTLEM_E_INVALID;
    if (!c->has_graph || !c->has_params) return fail(c, FASTLEM_E_STATE, "debug_fetch: no run yet");
    FL_CK(fl_set_device(c->device));
    const uint32_t n = c->n;
    if (stage == 9) {  // FL_FLOW_STATS counters (debug builds): 32 x u64, also resets them
        if (bytes != 256) return fail(c, FASTLEM_E_INVALID, "debug_fetch: flow stats are 256 bytes");
        FL_CK(fl_d2h(out, c->d_flow_stats, 256, c->stream));
        FL_CK(fl_memset(c->d_flow_stats, 0, 256, c->stream));
        FL_CK(fl_stream_sync(c->stream));
        return FASTLEM_OK;
    }
#ifdef FL_FLOW_STATS
    if (stage >= 100) {  // raw internal arrays in the CURRENT numbering (debug builds)
        const void* src = nullptr;
        size_t want = (size_t)n * 4;
        switch (stage) {
            case 100: src = c->d_tlog; want = (size_t)n * 32; break;
            case 101: src = L_(c).recv; break;
            case 102: src = c->d_sg_head; break;
            case 103: src = c->d_hgt; break;
            case 104: src = c->d_dirty_from; break;
            default: return fail(c, FASTLEM_E_INVALID, "debug_fetch: unknown raw stage");
        }
        if (bytes != want) return fail(c, FASTLEM_E_INVALID, "debug_fetch: wrong buffer size");
        FL_CK(fl_d2h(out, src, bytes, c->stream));
        if (stage == 100) FL_CK(fl_memset(c->d_tlog, 0, want, c->stream));
        FL_CK(fl_stream_sync(c->stream));
        return FASTLEM_OK;
    }
#endif
    if (!c->layout_valid && stage != FASTLEM_STAGE_FLOOD_RANK)
        return fail(c, FASTLEM_E_STATE, "debug_fetch: no run yet");
    Layout& L = L_(c);
    const bool is_f64 = stage == FASTLEM_STAGE_DRAINAGE_AREA || stage == FASTLEM_STAGE_RESPONSE_TIME ||
                        stage == FASTLEM_STAGE_ELEVATION;
    if (bytes != (size_t)n * (is_f64 ? 8 : 4)) return fail(c, FASTLEM_E_INVALID, "debug_fetch: wrong buffer size");
    const uint64_t launches_before = c->stats.kernel_launches;
    const void* src = nullptr;
    switch (stage) {
        case FASTLEM_STAGE_RECEIVERS:
            LAUNCH_N(k_unpermute_ids, n, n, L.orig_of, L.recv, c->d_out_u32);
            src = c->d_out_u32;
            break;
        case FASTLEM_STAGE_RECEIVERS_INITIAL:
            if (c->stages_valid) src = c->d_recv0;
            else { LAUNCH_N(k_unpermute_ids, n, n, L.orig_of, L.recv, c->d_out_u32); src = c->d_out_u32; }
            break;
        case FASTLEM_STAGE_LABELS_INITIAL:
            if (c->stages_valid) { src = c->d_label0; break; }
            // no lake removal happened: the initial labels are the final ones
            // fall through
        case FASTLEM_STAGE_LABELS:
        case FASTLEM_STAGE_DEPTH:
        case FASTLEM_STAGE_DRAINAGE_AREA:
        case FASTLEM_STAGE_RESPONSE_TIME: {
            // (root, depth) of the final forest, recomputed here so both sweeps share one definition
            FL_RC(run_labels(c));
            LAUNCH_N(k_labels_depth, n, n, c->d_pd, L.is_outlet, c->d_label, c->d_depth);
            if (stage == FASTLEM_STAGE_DEPTH) {
                LAUNCH_N(k_unpermute_u32, n, n, L.orig_of, c->d_depth, c->d_out_u32);
                src = c->d_out_u32;
            } else if (stage == FASTLEM_STAGE_DRAINAGE_AREA || stage == FASTLEM_STAGE_RESPONSE_TIME) {
                // sites outside every outlet's basin keep the loop's initial values (generator.rs:144-145)
                const bool area = stage == FASTLEM_STAGE_DRAINAGE_AREA;
                LAUNCH_N(k_stage_value, n, n, c->d_depth, area ? c->d_A : c->d_rt, area ? L.areas : nullptr,
                         c->d_out_f64 /*scratch*/);
                // d_out_f64 is in the current numbering here; unpermute through d_A-sized scratch d_rt? use d_pd as scratch
                LAUNCH_N(k_unpermute_f64, n, n, L.orig_of, c->d_out_f64, (double*)c->d_pd2);
                src = c->d_pd2;
            } else {
                LAUNCH_N(k_unpermute_ids, n, n, L.orig_of, c->d_label, c->d_out_u32);
                src = c->d_out_u32;
            }
            break;
        }
        case FASTLEM_STAGE_ELEVATION:
            LAUNCH_N(k_unpermute_f64, n, n, L.orig_of, L.elev, c->d_out_f64);
            src = c->d_out_f64;
            break;
        case FASTLEM_STAGE_FLOOD_RANK:
            FL_RC(ensure_rank(c));
            src = c->orig.rank;
            break;
        default: return fail(c, FASTLEM_E_INVALID, "debug_fetch: unknown stage");
    }
    FL_CK(fl_d2h(out, src, bytes, c->stream));
    FL_CK(fl_stream_sync(c->stream));
    FL_CK(fl_last_error());
    c->stats.kernel_launches = launches_before;  // debug traffic is not part of the run
    return FASTLEM_OK;
}

}  // extern "C"
