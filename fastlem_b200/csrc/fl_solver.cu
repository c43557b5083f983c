// fl_solver.cu -- host orchestration of the device-resident generate() loop + the C ABI
// (include/fastlem_b200.h).  Reference path: src/lem/generator.rs:90-213.
//
// Compiled by nvcc for sm_100a (product) or, with -DFL_EMU, by g++ as a serial host emulation used only
// by the CPU test tier (see fl_rt.h).
#include "../../include/fastlem_b200.h"

#include <chrono>
#include <cmath>
#include <new>
#include <string>
#include <vector>

#include "fl_flood.h"
#include "fl_kernels.cuh"

#ifdef FL_EMU
thread_local fl_dim3 threadIdx, blockIdx, blockDim, gridDim;
#endif

namespace {

double wall_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

enum Stage { ST_RECV = 0, ST_LABEL, ST_LAKE, ST_ORDER, ST_AREA, ST_ELEV, ST_COUNT };

}  // namespace

struct fastlem_ctx {
    int device = 0;
    cudaStream_t stream = 0;
    bool stream_ok = false;
    std::string err;

    // model (device)
    uint32_t n = 0, nnz = 0;
    bool has_graph = false, has_params = false;
    uint32_t* d_row_ptr = nullptr;
    uint32_t* d_col = nullptr;
    double* d_dist = nullptr;
    double* d_areas = nullptr;
    // borrowed host pointers (flood order is computed from them on first use)
    const uint32_t* h_row_ptr = nullptr;
    const uint32_t* h_col = nullptr;
    const double* h_dist = nullptr;

    // parameters (device)
    double* d_init = nullptr;
    double* d_erod = nullptr;
    double* d_uplift = nullptr;
    double* d_tan = nullptr;  // null = None everywhere
    uint8_t* d_is_outlet = nullptr;
    std::vector<uint32_t> outlets;

    // flood order (static per graph+outlets), lazily built
    bool rank_ready = false;
    uint32_t* d_rank = nullptr;
    uint32_t* d_rank_to_node = nullptr;

    // state
    double* d_elev = nullptr;
    uint32_t* d_recv = nullptr;
    double* d_drecv = nullptr;
    unsigned long long* d_pd = nullptr;
    unsigned long long* d_lake_key = nullptr;
    uint32_t* d_label = nullptr;
    uint32_t* d_depth = nullptr;
    uint32_t* d_ids = nullptr;
    uint32_t* d_sorted_depth = nullptr;
    uint32_t* d_order = nullptr;
    uint32_t* d_offs = nullptr;  // n+2
    double* d_A = nullptr;
    double* d_rt = nullptr;
    uint32_t* d_flags = nullptr;
    uint32_t* h_flags = nullptr;  // pinned
    uint32_t* h_offs = nullptr;   // pinned, n+2
    void* d_sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    // keep_stages
    uint32_t* d_recv0 = nullptr;
    uint32_t* d_label0 = nullptr;
    bool stages_valid = false;

    // options
    bool opt_profile = false, opt_keep = false;
    int64_t opt_sweep = 0;

    fastlem_stats stats{};
    cudaEvent_t ev[ST_COUNT + 1] = {};
    cudaEvent_t ev_run[2] = {};
    bool ev_ok = false;
};

namespace {

int fail(fastlem_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}

#define FL_CK(expr)                                                                                          \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return fail(c, FASTLEM_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));             \
    } while (0)

template <class T> void dfree(T*& p) {
    if (p) fl_free(p);
    p = nullptr;
}

template <class T> cudaError_t dalloc(T*& p, size_t count) {
    dfree(p);
    void* v = nullptr;
    cudaError_t e = fl_malloc(&v, count * sizeof(T));
    p = (T*)v;
    return e;
}

void free_graph(fastlem_ctx* c) {
    dfree(c->d_row_ptr); dfree(c->d_col); dfree(c->d_dist); dfree(c->d_areas);
    dfree(c->d_init); dfree(c->d_erod); dfree(c->d_uplift); dfree(c->d_tan); dfree(c->d_is_outlet);
    dfree(c->d_rank); dfree(c->d_rank_to_node);
    dfree(c->d_elev); dfree(c->d_recv); dfree(c->d_drecv); dfree(c->d_pd); dfree(c->d_lake_key);
    dfree(c->d_label); dfree(c->d_depth); dfree(c->d_ids); dfree(c->d_sorted_depth); dfree(c->d_order);
    dfree(c->d_offs); dfree(c->d_A); dfree(c->d_rt); dfree(c->d_recv0); dfree(c->d_label0);
    if (c->d_sort_tmp) fl_free(c->d_sort_tmp);
    c->d_sort_tmp = nullptr;
    if (c->h_offs) fl_free_host(c->h_offs);
    c->h_offs = nullptr;
    c->has_graph = c->has_params = c->rank_ready = c->stages_valid = false;
}

inline unsigned blocks_for(uint32_t count) { return (count + 255u) / 256u; }

// lazily compute + upload the flood order (fl_flood.cpp)
int ensure_rank(fastlem_ctx* c) {
    if (c->rank_ready) return FASTLEM_OK;
    double t0 = wall_ms();
    const uint32_t n = c->n;
    std::vector<uint32_t> rank(n), inv(n, FL_NONE);
    fl_flood_rank(n, c->h_row_ptr, c->h_col, c->h_dist, c->outlets.data(), (uint32_t)c->outlets.size(), rank.data());
    for (uint32_t i = 0; i < n; ++i)
        if (rank[i] != FL_NONE) inv[rank[i]] = i;
    FL_CK(dalloc(c->d_rank, n));
    FL_CK(dalloc(c->d_rank_to_node, n));
    FL_CK(fl_h2d(c->d_rank, rank.data(), sizeof(uint32_t) * n, c->stream));
    FL_CK(fl_h2d(c->d_rank_to_node, inv.data(), sizeof(uint32_t) * n, c->stream));
    FL_CK(fl_stream_sync(c->stream));
    c->rank_ready = true;
    c->stats.ms_flood_rank = wall_ms() - t0;
    return FASTLEM_OK;
}

// K2 driver: pointer jumping until stable; leaves (root, depth) pairs in d_pd
int run_jump(fastlem_ctx* c) {
    const uint32_t n = c->n;
    const unsigned g = blocks_for(n);
    FL_LAUNCH(k_jump_init, g, 256, c->stream, n, c->d_recv, c->d_pd);
    c->stats.kernel_launches++; c->stats.n_labels++;
    for (int round = 0; round < 40; ++round) {
        FL_CK(fl_memset(c->d_flags + FL_FLAG_JUMP, 0, sizeof(uint32_t), c->stream));
        FL_LAUNCH(k_jump, g, 256, c->stream, n, c->d_pd, c->d_flags);
        c->stats.kernel_launches++; c->stats.n_labels++;
        FL_CK(fl_d2h(c->h_flags, c->d_flags, sizeof(uint32_t) * FL_N_FLAGS, c->stream));
        FL_CK(fl_stream_sync(c->stream));
        if (!c->h_flags[FL_FLAG_JUMP]) return FASTLEM_OK;
    }
    return fail(c, FASTLEM_E_STATE, "pointer jumping did not converge (cycle in receivers?)");
}

int stage_mark(fastlem_ctx* c, int k) {
    if (c->opt_profile) FL_CK(fl_event_record(c->ev[k], c->stream));
    return FASTLEM_OK;
}

// one loop body of generator.rs:140-210; *changed_out = the `changed` flag
int iterate(fastlem_ctx* c, bool first, bool* changed_out) {
    const uint32_t n = c->n;
    const unsigned g = blocks_for(n);
    int rc;
    FL_CK(fl_memset(c->d_flags, 0, sizeof(uint32_t) * FL_N_FLAGS, c->stream));
    if ((rc = stage_mark(c, 0))) return rc;

    // K1 receivers
    FL_LAUNCH(k_receivers, g, 256, c->stream, n, c->d_row_ptr, c->d_col, c->d_dist, c->d_elev, c->d_is_outlet,
              c->d_recv, c->d_drecv, c->d_flags);
    c->stats.kernel_launches++; c->stats.n_receivers++;
    if ((rc = stage_mark(c, 1))) return rc;

    // K2 labels (+ depth)
    if ((rc = run_jump(c))) return rc;
    const bool has_lake = c->h_flags[FL_FLAG_LAKE] != 0;
    if ((rc = stage_mark(c, 2))) return rc;

    // K3 lake connection
    c->stages_valid = false;
    if (has_lake) {
        if ((rc = ensure_rank(c))) return rc;
        FL_LAUNCH(k_labels_only, g, 256, c->stream, n, c->d_pd, c->d_label);
        c->stats.kernel_launches++;
        if (c->opt_keep) {
            FL_CK(fl_d2d(c->d_recv0, c->d_recv, sizeof(uint32_t) * n, c->stream));
            FL_CK(fl_d2d(c->d_label0, c->d_label, sizeof(uint32_t) * n, c->stream));
            c->stages_valid = true;
        }
        FL_CK(fl_memset(c->d_lake_key, 0xFF, sizeof(unsigned long long) * n, c->stream));
        FL_LAUNCH(k_lake_min, g, 256, c->stream, n, c->d_row_ptr, c->d_col, c->d_label, c->d_is_outlet, c->d_rank,
                  c->d_lake_key);
        FL_LAUNCH(k_lake_reverse, g, 256, c->stream, n, c->d_row_ptr, c->d_col, c->d_dist, c->d_is_outlet,
                  c->d_rank_to_node, c->d_lake_key, c->d_label, c->d_recv, c->d_drecv);
        c->stats.kernel_launches += 2; c->stats.n_lakes += 3;
        c->stats.lake_iterations++;
        if ((rc = run_jump(c))) return rc;
    }
    if ((rc = stage_mark(c, 3))) return rc;

    // ordering: sort nodes by depth
    FL_LAUNCH(k_labels_finalize, g, 256, c->stream, n, c->d_pd, c->d_is_outlet, c->d_areas, c->d_label, c->d_depth,
              c->d_ids, c->d_A, c->d_rt);
    FL_CK(fl_sort_pairs(c->d_sort_tmp, c->sort_tmp_bytes, c->d_depth, c->d_sorted_depth, c->d_ids, c->d_order, n, 32,
                        c->stream, false));
    FL_CK(fl_memset(c->d_flags + FL_FLAG_MAXDEPTH, 0xFF, sizeof(uint32_t), c->stream));
    FL_LAUNCH(k_level_offsets, g, 256, c->stream, n, c->d_sorted_depth, c->d_offs, c->d_flags);
    c->stats.kernel_launches += 3; c->stats.n_order += 3;
    FL_CK(fl_d2h(c->h_flags, c->d_flags, sizeof(uint32_t) * FL_N_FLAGS, c->stream));
    FL_CK(fl_stream_sync(c->stream));
    const uint32_t maxd = c->h_flags[FL_FLAG_MAXDEPTH];
    if ((rc = stage_mark(c, 4))) return rc;
    if (maxd != FL_NONE) {
        FL_CK(fl_d2h(c->h_offs, c->d_offs, sizeof(uint32_t) * ((size_t)maxd + 2), c->stream));
        FL_CK(fl_stream_sync(c->stream));
        if (first) c->stats.depth_first = maxd + 1;
        c->stats.depth_last = maxd + 1;

        // K4 drainage area: deepest level first
        for (uint32_t lv = maxd + 1; lv-- > 0;) {
            const uint32_t b = c->h_offs[lv], cnt = c->h_offs[lv + 1] - b;
            FL_LAUNCH(k_area_level, blocks_for(cnt), 256, c->stream, b, cnt, c->d_order, c->d_row_ptr, c->d_col,
                      c->d_recv, c->d_areas, c->d_A);
        }
        c->stats.kernel_launches += maxd + 1; c->stats.n_area += maxd + 1;
        if ((rc = stage_mark(c, 5))) return rc;

        // K5 response time + elevation: roots first
        for (uint32_t lv = 0; lv <= maxd; ++lv) {
            const uint32_t b = c->h_offs[lv], cnt = c->h_offs[lv + 1] - b;
            FL_LAUNCH(k_elev_level, blocks_for(cnt), 256, c->stream, b, cnt, (int)lv, c->d_order, c->d_recv, c->d_label,
                      c->d_drecv, c->d_A, c->d_erod, c->d_uplift, c->d_tan, c->d_elev, c->d_rt, c->d_flags);
        }
        c->stats.kernel_launches += maxd + 1; c->stats.n_elevation += maxd + 1;
    } else {
        if ((rc = stage_mark(c, 5))) return rc;
    }
    if ((rc = stage_mark(c, 6))) return rc;
    FL_CK(fl_d2h(c->h_flags, c->d_flags, sizeof(uint32_t) * FL_N_FLAGS, c->stream));
    FL_CK(fl_stream_sync(c->stream));
    FL_CK(fl_last_error());
    *changed_out = c->h_flags[FL_FLAG_CHANGED] != 0;

    if (c->opt_profile) {
        double* acc[ST_COUNT] = {&c->stats.ms_receivers, &c->stats.ms_labels, &c->stats.ms_lakes,
                                 &c->stats.ms_order,     &c->stats.ms_area,   &c->stats.ms_elevation};
        for (int k = 0; k < ST_COUNT; ++k) {
            float ms = 0.f;
            FL_CK(fl_event_elapsed(&ms, c->ev[k], c->ev[k + 1]));
            *acc[k] += ms;
        }
    }
    return FASTLEM_OK;
}

}  // namespace

extern "C" {

const char* fastlem_version(void) {
#ifdef FL_EMU
    return "fastlem_b200 0.1.0 emu";
#else
    return "fastlem_b200 0.1.0 sm_100a";
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
    void* hf = nullptr;
    if (fl_malloc_host(&hf, sizeof(uint32_t) * FL_N_FLAGS) != cudaSuccess) {
        fl_stream_destroy(c->stream);
        delete c;
        return FASTLEM_E_NOMEM;
    }
    c->h_flags = (uint32_t*)hf;
    bool ok = true;
    for (int k = 0; k <= ST_COUNT; ++k) ok = ok && fl_event_create(&c->ev[k]) == cudaSuccess;
    for (int k = 0; k < 2; ++k) ok = ok && fl_event_create(&c->ev_run[k]) == cudaSuccess;
    c->ev_ok = ok;
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
    free_graph(c);
    dfree(c->d_flags);
    if (c->h_flags) fl_free_host(c->h_flags);
    for (int k = 0; k <= ST_COUNT; ++k)
        if (c->ev[k]) fl_event_destroy(c->ev[k]);
    for (int k = 0; k < 2; ++k)
        if (c->ev_run[k]) fl_event_destroy(c->ev_run[k]);
    if (c->stream_ok) fl_stream_destroy(c->stream);
    delete c;
}

const char* fastlem_last_error(const fastlem_ctx* c) { return c ? c->err.c_str() : "null context"; }

int fastlem_set_option(fastlem_ctx* c, const char* name, int64_t value) {
    if (!c || !name) return FASTLEM_E_INVALID;
    std::string s(name);
    if (s == "profile") c->opt_profile = value != 0;
    else if (s == "keep_stages") c->opt_keep = value != 0;
    else if (s == "sweep") c->opt_sweep = value;
    else return fail(c, FASTLEM_E_INVALID, "unknown option: " + s);
    return FASTLEM_OK;
}

int fastlem_set_graph(fastlem_ctx* c, uint32_t n, const uint32_t* row_ptr, const uint32_t* col, const double* dist,
                      const double* areas) {
    if (!c) return FASTLEM_E_INVALID;
    if (!row_ptr || !areas) return fail(c, FASTLEM_E_INVALID, "set_graph: null pointer");
    if (n == 0 || n >= FL_NONE) return fail(c, FASTLEM_E_INVALID, "set_graph: n must be in [1, 2^32-2]");
    if (row_ptr[0] != 0) return fail(c, FASTLEM_E_INVALID, "set_graph: row_ptr[0] must be 0");
    for (uint32_t i = 0; i < n; ++i)
        if (row_ptr[i + 1] < row_ptr[i]) return fail(c, FASTLEM_E_INVALID, "set_graph: row_ptr must be non-decreasing");
    const uint32_t nnz = row_ptr[n];
    if (nnz && (!col || !dist)) return fail(c, FASTLEM_E_INVALID, "set_graph: null col/dist");
    for (uint32_t s = 0; s < nnz; ++s)
        if (col[s] >= n) return fail(c, FASTLEM_E_INVALID, "set_graph: neighbour index out of range");
    FL_CK(fl_set_device(c->device));
    double t0 = wall_ms();
    free_graph(c);
    c->n = n;
    c->nnz = nnz;
    c->h_row_ptr = row_ptr; c->h_col = col; c->h_dist = dist;
    FL_CK(dalloc(c->d_row_ptr, (size_t)n + 1));
    FL_CK(dalloc(c->d_col, nnz));
    FL_CK(dalloc(c->d_dist, nnz));
    FL_CK(dalloc(c->d_areas, n));
    FL_CK(fl_h2d(c->d_row_ptr, row_ptr, sizeof(uint32_t) * ((size_t)n + 1), c->stream));
    if (nnz) {
        FL_CK(fl_h2d(c->d_col, col, sizeof(uint32_t) * nnz, c->stream));
        FL_CK(fl_h2d(c->d_dist, dist, sizeof(double) * nnz, c->stream));
    }
    FL_CK(fl_h2d(c->d_areas, areas, sizeof(double) * n, c->stream));
    // state buffers
    FL_CK(dalloc(c->d_elev, n));
    FL_CK(dalloc(c->d_recv, n));
    FL_CK(dalloc(c->d_drecv, n));
    FL_CK(dalloc(c->d_pd, n));
    FL_CK(dalloc(c->d_lake_key, n));
    FL_CK(dalloc(c->d_label, n));
    FL_CK(dalloc(c->d_depth, n));
    FL_CK(dalloc(c->d_ids, n));
    FL_CK(dalloc(c->d_sorted_depth, n));
    FL_CK(dalloc(c->d_order, n));
    FL_CK(dalloc(c->d_offs, (size_t)n + 2));
    FL_CK(dalloc(c->d_A, n));
    FL_CK(dalloc(c->d_rt, n));
    FL_CK(dalloc(c->d_recv0, n));
    FL_CK(dalloc(c->d_label0, n));
    if (!c->d_flags) FL_CK(dalloc(c->d_flags, FL_N_FLAGS));
    void* ho = nullptr;
    FL_CK(fl_malloc_host(&ho, sizeof(uint32_t) * ((size_t)n + 2)));
    c->h_offs = (uint32_t*)ho;
    c->sort_tmp_bytes = 0;
    FL_CK(fl_sort_pairs(nullptr, c->sort_tmp_bytes, c->d_depth, c->d_sorted_depth, c->d_ids, c->d_order, n, 32,
                        c->stream, true));
    FL_CK(fl_malloc(&c->d_sort_tmp, c->sort_tmp_bytes));
    FL_CK(fl_stream_sync(c->stream));
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
    for (uint32_t k = 0; k < n_outlets; ++k)
        if (outlets[k] >= n) return fail(c, FASTLEM_E_INVALID, "set_parameters: outlet index out of range");
    FL_CK(fl_set_device(c->device));
    double t0 = wall_ms();
    FL_CK(dalloc(c->d_init, n));
    FL_CK(dalloc(c->d_erod, n));
    FL_CK(dalloc(c->d_uplift, n));
    FL_CK(dalloc(c->d_is_outlet, n));
    FL_CK(fl_h2d(c->d_init, initial_elevation, sizeof(double) * n, c->stream));
    FL_CK(fl_h2d(c->d_erod, erodibility, sizeof(double) * n, c->stream));
    FL_CK(fl_h2d(c->d_uplift, uplift_rate, sizeof(double) * n, c->stream));
    if (tan_max_slope) {
        FL_CK(dalloc(c->d_tan, n));
        FL_CK(fl_h2d(c->d_tan, tan_max_slope, sizeof(double) * n, c->stream));
    } else {
        dfree(c->d_tan);
    }
    // stream_tree.rs:101-107 outlet table (static across iterations, so built once)
    std::vector<uint8_t> table(n, 0);
    for (uint32_t k = 0; k < n_outlets; ++k) table[outlets[k]] = 1;
    FL_CK(fl_h2d(c->d_is_outlet, table.data(), n, c->stream));
    FL_CK(fl_stream_sync(c->stream));
    c->outlets.assign(outlets, outlets + n_outlets);
    c->rank_ready = false;
    c->has_params = true;
    c->stats.ms_upload += wall_ms() - t0;
    return FASTLEM_OK;
}

int fastlem_run(fastlem_ctx* c, uint32_t max_iteration, uint32_t* iterations_done) {
    if (!c) return FASTLEM_E_INVALID;
    if (!c->has_graph) return fail(c, FASTLEM_E_STATE, "run: model not set (ModelNotSet)");
    if (!c->has_params) return fail(c, FASTLEM_E_STATE, "run: parameters not set (ParametersNotSet)");
    FL_CK(fl_set_device(c->device));
    const uint32_t n = c->n;
    // reset per-run stats, keep the one-off ones
    double up = c->stats.ms_upload, fr = c->stats.ms_flood_rank;
    c->stats = fastlem_stats{};
    c->stats.ms_upload = up;
    c->stats.ms_flood_rank = fr;
    FL_CK(fl_event_record(c->ev_run[0], c->stream));
    FL_CK(fl_d2d(c->d_elev, c->d_init, sizeof(double) * n, c->stream));
    uint32_t it = 0;
    while (it < max_iteration) {
        bool changed = false;
        int rc = iterate(c, it == 0, &changed);
        if (rc) return rc;
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

int fastlem_download(fastlem_ctx* c, double* elevations_out) {
    if (!c || !elevations_out) return FASTLEM_E_INVALID;
    if (!c->has_graph || !c->has_params) return fail(c, FASTLEM_E_STATE, "download: nothing to download");
    FL_CK(fl_set_device(c->device));
    double t0 = wall_ms();
    FL_CK(fl_d2h(elevations_out, c->d_elev, sizeof(double) * c->n, c->stream));
    FL_CK(fl_stream_sync(c->stream));
    c->stats.ms_download = wall_ms() - t0;
    return FASTLEM_OK;
}

int fastlem_download_to_device(fastlem_ctx* c, double* device_out) {
    if (!c || !device_out) return FASTLEM_E_INVALID;
    if (!c->has_graph || !c->has_params) return fail(c, FASTLEM_E_STATE, "download: nothing to download");
    FL_CK(fl_set_device(c->device));
    FL_CK(fl_d2d(device_out, c->d_elev, sizeof(double) * c->n, c->stream));
    FL_CK(fl_stream_sync(c->stream));
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
    const size_t n = c->n;
    const void* src = nullptr;
    size_t want = 0;
    switch (stage) {
        case FASTLEM_STAGE_RECEIVERS: src = c->d_recv; want = n * 4; break;
        case FASTLEM_STAGE_RECEIVERS_INITIAL: src = c->stages_valid ? c->d_recv0 : c->d_recv; want = n * 4; break;
        case FASTLEM_STAGE_LABELS_INITIAL: src = c->stages_valid ? c->d_label0 : c->d_label; want = n * 4; break;
        case FASTLEM_STAGE_LABELS: src = c->d_label; want = n * 4; break;
        case FASTLEM_STAGE_DEPTH: src = c->d_depth; want = n * 4; break;
        case FASTLEM_STAGE_DRAINAGE_AREA: src = c->d_A; want = n * 8; break;
        case FASTLEM_STAGE_RESPONSE_TIME: src = c->d_rt; want = n * 8; break;
        case FASTLEM_STAGE_ELEVATION: src = c->d_elev; want = n * 8; break;
        case FASTLEM_STAGE_FLOOD_RANK: {
            int rc = ensure_rank(c);
            if (rc) return rc;
            src = c->d_rank; want = n * 4; break;
        }
        default: return fail(c, FASTLEM_E_INVALID, "debug_fetch: unknown stage");
    }
    if (bytes != want) return fail(c, FASTLEM_E_INVALID, "debug_fetch: wrong buffer size");
    FL_CK(fl_d2h(out, src, want, c->stream));
    FL_CK(fl_stream_sync(c->stream));
    return FASTLEM_OK;
}

}  // extern "C"
