// fl_interp.cu -- host side + C ABI of the natural-neighbour interpolator (include/fastlem_b200.h, section
// "Terrain2D::get_elevation").  Mirrors TerrainInterpolator2D (reference src/models/surface/interpolator.rs:6-28):
// `new(sites)` -> fastlem_interp_create (triangulation handed over by the caller, who already has delaunator's),
// `interpolate(elevations, site)` -> fastlem_interp_set_values + fastlem_interp_points / fastlem_interp_raster.
//
// Compiled by nvcc for sm_100a (product) or, with -DFL_EMU, by g++ as the serial host emulation of the CPU test
// tier.  There is no CPU implementation behind the product entry points.
#include "../../include/fastlem_b200.h"

#include <chrono>
#include <cmath>
#include <new>
#include <string>

#include "fl_interp.cuh"

struct fastlem_interp {
    int device = 0;
    cudaStream_t stream{};
    bool stream_ok = false;
    std::string err;
    uint32_t n_sites = 0, n_tri = 0;
    FliPt* d_site = nullptr;
    FliTri* d_tri = nullptr;
    FliNbr* d_nbr = nullptr;
    FliCirc* d_circ = nullptr;
    FliGeo* d_geo = nullptr;  // FLI_DENORM builds only
    uint32_t* d_cell = nullptr;
    uint32_t* d_cell2 = nullptr;
    double* d_value = nullptr;
    uint32_t* d_flags = nullptr;
    uint32_t* h_flags = nullptr;  // pinned
    double* d_out = nullptr;      // result staging for the host-pointer entry points
    size_t out_cap = 0;
    FliPt* d_query = nullptr;
    size_t query_cap = 0;
    bool has_values = false;
    FliGrid grid{};
    double sgn = 1.0;
    double origin[2] = {0.0, 0.0};  // lower corner of the sites' bounding box, subtracted from every coordinate
    uint32_t max_walk = 0;
    cudaEvent_t ev[2] = {};
    fastlem_interp_stats stats{};
};

namespace {

double wall_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int fail(fastlem_interp* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}

#define FLI_CK(expr)                                                                             \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return fail(c, FASTLEM_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

template <class T> cudaError_t ialloc(T*& p, size_t count) {
    void* v = nullptr;
    cudaError_t e = fl_malloc(&v, (count ? count : 1) * sizeof(T));
    p = (T*)v;
    return e;
}

inline unsigned blocks_for(size_t count, unsigned block = 256) { return (unsigned)((count + block - 1u) / block); }

#define FLI_LAUNCH_N(kernel, count, ...)                                       \
    do {                                                                       \
        if ((count) > 0) {                                                     \
            FL_LAUNCH(kernel, blocks_for(count), 256, c->stream, __VA_ARGS__); \
            c->stats.kernel_launches++;                                        \
        }                                                                      \
    } while (0)

int read_flags(fastlem_interp* c) {
    FLI_CK(fl_d2h(c->h_flags, c->d_flags, sizeof(uint32_t) * FLI_N_FLAGS, c->stream));
    FLI_CK(fl_stream_sync(c->stream));
    FLI_CK(fl_last_error());
    return FASTLEM_OK;
}

FliModel model_of(const fastlem_interp* c) {
    FliModel M;
    M.site = c->d_site;
    M.tri = c->d_tri;
    M.nbr = c->d_nbr;
    M.circ = c->d_circ;
    M.geo = c->d_geo;
    M.cell = c->d_cell;
    M.value = c->d_value;
    M.grid = c->grid;
    M.n_tri = c->n_tri;
    M.max_walk = c->max_walk;
    M.sgn = c->sgn;
    M.ox = c->origin[0];
    M.oy = c->origin[1];
    return M;
}

int check_query_flags(fastlem_interp* c) {
    int rc = read_flags(c);
    if (rc) return rc;
    if (c->h_flags[FLI_F_WALK_OVERFLOW])
        return fail(c, FASTLEM_E_INVALID, "interpolate: point location did not terminate (inconsistent triangulation)");
    if (c->h_flags[FLI_F_CAVITY_OVERFLOW])
        return fail(c, FASTLEM_E_INVALID, "interpolate: a natural-neighbour cavity exceeded the supported size "
                                          "(degenerate or non-Delaunay triangulation)");
    return FASTLEM_OK;
}

int ensure_out(fastlem_interp* c, size_t count) {
    if (count <= c->out_cap) return FASTLEM_OK;
    if (c->d_out) fl_free(c->d_out, c->stream);
    c->d_out = nullptr;
    c->out_cap = 0;
    FLI_CK(ialloc(c->d_out, count));
    c->out_cap = count;
    return FASTLEM_OK;
}

int validate_raster(fastlem_interp* c, const fastlem_raster* r) {
    if (!r) return fail(c, FASTLEM_E_INVALID, "raster: null descriptor");
    if (r->width == 0 || r->height == 0) return fail(c, FASTLEM_E_INVALID, "raster: empty image");
    if (r->row_begin > r->row_end || r->row_end > r->height)
        return fail(c, FASTLEM_E_INVALID, "raster: row range outside the image");
    return FASTLEM_OK;
}

int launch_raster(fastlem_interp* c, const fastlem_raster* r, double* d_out) {
    FliRaster R;
    R.x0 = r->x0; R.y0 = r->y0; R.span_x = r->span_x; R.span_y = r->span_y; R.offset = r->pixel_offset;
    R.width = r->width; R.height = r->height; R.row_begin = r->row_begin; R.row_end = r->row_end;
    const uint32_t rows = r->row_end - r->row_begin;
    const uint64_t tiles64 = (uint64_t)((r->width + 15u) / 16u) * (uint64_t)((rows + 15u) / 16u);
    if (tiles64 > 0x7FFFFFFFull)
        return fail(c, FASTLEM_E_INVALID, "raster: too many pixels for one call, split the row range");
    const uint32_t tiles = (uint32_t)tiles64;
    FLI_CK(fl_memset(c->d_flags, 0, sizeof(uint32_t) * FLI_N_FLAGS, c->stream));
    FLI_CK(fl_event_record(c->ev[0], c->stream));
    if (tiles) {
        FL_LAUNCH(k_nn_raster, tiles, 256, c->stream, model_of(c), R, d_out, c->d_flags);
        c->stats.kernel_launches++;
    }
    FLI_CK(fl_event_record(c->ev[1], c->stream));
    return FASTLEM_OK;
}

int finish_query(fastlem_interp* c, uint64_t queries) {
    int rc = check_query_flags(c);  // synchronises the stream
    if (rc) return rc;
    float ms = 0.f;
    FLI_CK(fl_event_elapsed(&ms, c->ev[0], c->ev[1]));
    c->stats.ms_query_kernel = ms;
    c->stats.queries = queries;
    return FASTLEM_OK;
}

}  // namespace

extern "C" {

void fastlem_interp_destroy(fastlem_interp* c) {
    if (!c) return;
    fl_set_device(c->device);
    void* ptrs[] = {c->d_site, c->d_tri, c->d_nbr, c->d_circ, c->d_geo, c->d_cell, c->d_cell2, c->d_value, c->d_flags, c->d_out,
                    c->d_query};
    for (void* p : ptrs)
        if (p) fl_free(p, c->stream);
    if (c->h_flags) fl_free_host(c->h_flags);
    for (int k = 0; k < 2; ++k)
        if (c->ev[k]) fl_event_destroy(c->ev[k]);
    if (c->stream_ok) fl_stream_destroy(c->stream);
    delete c;
}

const char* fastlem_interp_last_error(const fastlem_interp* c) { return c ? c->err.c_str() : "null interpolator"; }

static int interp_setup(fastlem_interp* c, const double* sites_xy, const uint32_t* triangles, const uint32_t* halfedges) {
    const uint32_t n = c->n_sites, nt = c->n_tri;
    const double t0 = wall_ms();
    // bounding box of the sites; its lower corner becomes the origin of all device-side coordinates, so that data given
    // in large absolute coordinates (map projections) keeps its precision in the circumcentre / area arithmetic
    double lo[2] = {INFINITY, INFINITY}, hi[2] = {-INFINITY, -INFINITY};
    for (uint32_t i = 0; i < n; ++i)
        for (int k = 0; k < 2; ++k) {
            const double v = sites_xy[2 * (size_t)i + k];
            if (!(v == v) || std::isinf(v)) return fail(c, FASTLEM_E_INVALID, "interpolator: non-finite site coordinate");
            lo[k] = v < lo[k] ? v : lo[k];
            hi[k] = v > hi[k] ? v : hi[k];
        }
    if (n == 0) lo[0] = lo[1] = hi[0] = hi[1] = 0.0;
    c->origin[0] = lo[0];
    c->origin[1] = lo[1];
    void* hf = nullptr;
    FLI_CK(fl_malloc_host(&hf, sizeof(uint32_t) * FLI_N_FLAGS));
    c->h_flags = (uint32_t*)hf;
    for (int k = 0; k < 2; ++k) FLI_CK(fl_event_create(&c->ev[k]));
    FLI_CK(ialloc(c->d_flags, FLI_N_FLAGS));
    FLI_CK(ialloc(c->d_site, n));
    FLI_CK(ialloc(c->d_tri, nt));
    FLI_CK(ialloc(c->d_nbr, nt));
    FLI_CK(ialloc(c->d_circ, nt));
#if FLI_DENORM
    FLI_CK(ialloc(c->d_geo, nt));
#endif
    FLI_CK(ialloc(c->d_value, n));
    // the raw delaunator arrays are only needed by k_nn_prepare
    uint32_t *d_triangles = nullptr, *d_halfedges = nullptr;
    FLI_CK(ialloc(d_triangles, 3 * (size_t)nt));
    cudaError_t e = ialloc(d_halfedges, 3 * (size_t)nt);
    if (e != cudaSuccess) { fl_free(d_triangles, c->stream); FLI_CK(e); }
    int rc = FASTLEM_OK;
    do {
        e = fl_h2d(c->d_site, sites_xy, sizeof(double) * 2 * (size_t)n, c->stream);
        if (e == cudaSuccess) e = fl_h2d(d_triangles, triangles, sizeof(uint32_t) * 3 * (size_t)nt, c->stream);
        if (e == cudaSuccess) e = fl_h2d(d_halfedges, halfedges, sizeof(uint32_t) * 3 * (size_t)nt, c->stream);
        if (e == cudaSuccess) e = fl_memset(c->d_flags, 0, sizeof(uint32_t) * FLI_N_FLAGS, c->stream);
        if (e != cudaSuccess) { rc = fail(c, FASTLEM_E_CUDA, std::string("upload: ") + cudaGetErrorString(e)); break; }
        if (n) {
            FL_LAUNCH(k_nn_translate, blocks_for(n), 256, c->stream, n, c->d_site, c->origin[0], c->origin[1]);
            c->stats.kernel_launches++;
        }
        if (nt) {
            FL_LAUNCH(k_nn_prepare, blocks_for(nt), 256, c->stream, n, nt, c->d_site, d_triangles, d_halfedges, c->d_tri,
                      c->d_nbr, c->d_circ, c->d_geo, c->d_flags);
            c->stats.kernel_launches++;
        }
        rc = read_flags(c);
    } while (0);
    fl_free(d_triangles, c->stream);
    fl_free(d_halfedges, c->stream);
    if (rc) return rc;
    const uint32_t* f = c->h_flags;
    if (f[FLI_F_BAD_INDEX]) return fail(c, FASTLEM_E_INVALID, "interpolator: triangle vertex index out of range");
    if (f[FLI_F_BAD_HALFEDGE])
        return fail(c, FASTLEM_E_INVALID, "interpolator: halfedges do not pair opposite edges of the triangles");
    if (f[FLI_F_DEGENERATE])
        return fail(c, FASTLEM_E_INVALID, "interpolator: degenerate (zero-area or non-finite) triangle");
    if (f[FLI_F_POS] && f[FLI_F_NEG])
        return fail(c, FASTLEM_E_INVALID, "interpolator: triangles are not consistently oriented");
    c->sgn = f[FLI_F_NEG] ? -1.0 : 1.0;
    c->stats.clockwise = f[FLI_F_NEG] ? 1u : 0u;
    FLI_LAUNCH_N(k_nn_check_delaunay, nt, nt, c->d_site, c->d_tri, c->d_nbr, c->d_circ, c->d_flags);
    rc = read_flags(c);
    if (rc) return rc;
    if (c->h_flags[FLI_F_NOT_DELAUNAY])
        return fail(c, FASTLEM_E_INVALID, "interpolator: not a Delaunay triangulation (" +
                                              std::to_string(c->h_flags[FLI_F_NOT_DELAUNAY]) +
                                              " opposite vertices inside a circumcircle)");

    // hint grid over the bounding box of the (translated) sites: ~2 triangle centroids per cell
    const double w = hi[0] - lo[0], h = hi[1] - lo[1];
    double cells = nt / 2.0;
    if (cells < 1.0) cells = 1.0;
    if (cells > 64.0e6) cells = 64.0e6;
    double gx = 1.0, gy = 1.0;
    if (w > 0.0 && h > 0.0) {
        gx = std::ceil(std::sqrt(cells * w / h));
        gy = std::ceil(cells / gx);
    }
    if (gx < 1.0) gx = 1.0;
    if (gy < 1.0) gy = 1.0;
    if (gx > 16384.0) gx = 16384.0;
    if (gy > 16384.0) gy = 16384.0;
    c->grid.gx = (uint32_t)gx;
    c->grid.gy = (uint32_t)gy;
    c->grid.x0 = 0.0;  // device coordinates are relative to the lower corner of the bounding box
    c->grid.y0 = 0.0;
    c->grid.inv_cell_x = w > 0.0 ? gx / w : 0.0;
    c->grid.inv_cell_y = h > 0.0 ? gy / h : 0.0;
    const size_t nc = (size_t)c->grid.gx * c->grid.gy;
    c->stats.grid_x = c->grid.gx;
    c->stats.grid_y = c->grid.gy;
    // a walk across the whole domain crosses O(gx + gy) triangles on a uniform triangulation; generous cap
    c->max_walk = 64u * (c->grid.gx + c->grid.gy) + 4096u;
    FLI_CK(ialloc(c->d_cell, nc));
    FLI_CK(ialloc(c->d_cell2, nc));
    FLI_CK(fl_memset(c->d_cell, 0xFF, sizeof(uint32_t) * nc, c->stream));
    FLI_LAUNCH_N(k_nn_grid_fill, nt, nt, c->d_site, c->d_tri, c->grid, c->d_cell);
    uint32_t passes = 0;
    while (nt) {
        FLI_CK(fl_memset(c->d_flags, 0, sizeof(uint32_t) * FLI_N_FLAGS, c->stream));
        FLI_LAUNCH_N(k_nn_grid_dilate, nc, c->grid, c->d_cell, c->d_cell2, c->d_flags);
        uint32_t* t = c->d_cell; c->d_cell = c->d_cell2; c->d_cell2 = t;
        rc = read_flags(c);
        if (rc) return rc;
        ++passes;
        if (c->h_flags[FLI_F_EMPTY_CELLS] == 0) break;
        if (passes > c->grid.gx + c->grid.gy) return fail(c, FASTLEM_E_INVALID, "interpolator: hint grid could not be filled");
    }
    c->stats.grid_passes = passes;
    fl_free(c->d_cell2, c->stream);
    c->d_cell2 = nullptr;
    c->stats.ms_setup = wall_ms() - t0;
    return FASTLEM_OK;
}

int fastlem_interp_create(fastlem_interp** out, int device_ordinal, uint32_t n_sites, const double* sites_xy,
                          uint32_t n_triangles, const uint32_t* triangles, const uint32_t* halfedges) {
    if (!out) return FASTLEM_E_INVALID;
    *out = nullptr;
    if ((n_sites && !sites_xy) || (n_triangles && (!triangles || !halfedges))) return FASTLEM_E_INVALID;
    if (n_triangles > 0x55555554u) return FASTLEM_E_INVALID;  // 3T must fit the uint32 half-edge ids
    fastlem_interp* c = new (std::nothrow) fastlem_interp();
    if (!c) return FASTLEM_E_NOMEM;
    c->device = device_ordinal;
    c->n_sites = n_sites;
    c->n_tri = n_triangles;
    cudaError_t e = fl_set_device(device_ordinal);
    if (e == cudaSuccess) e = fl_stream_create(&c->stream);
    if (e != cudaSuccess) {
        // no CPU fallback: without a usable CUDA device there is no interpolator
        std::fprintf(stderr, "fastlem_b200: cannot create interpolator on CUDA device %d: %s\n", device_ordinal,
                     cudaGetErrorString(e));
        delete c;
        return FASTLEM_E_CUDA;
    }
    c->stream_ok = true;
    int rc = interp_setup(c, sites_xy, triangles, halfedges);
    if (rc) {
        std::fprintf(stderr, "fastlem_b200: fastlem_interp_create: %s\n", c->err.c_str());
        fastlem_interp_destroy(c);
        return rc;
    }
    *out = c;
    return FASTLEM_OK;
}

int fastlem_interp_set_values(fastlem_interp* c, const double* values) {
    if (!c || (!values && c->n_sites)) return FASTLEM_E_INVALID;
    FLI_CK(fl_set_device(c->device));
    FLI_CK(fl_h2d(c->d_value, values, sizeof(double) * c->n_sites, c->stream));
    FLI_CK(fl_stream_sync(c->stream));
    c->has_values = true;
    return FASTLEM_OK;
}

int fastlem_interp_set_values_device(fastlem_interp* c, const double* device_values) {
    if (!c || (!device_values && c->n_sites)) return FASTLEM_E_INVALID;
    FLI_CK(fl_set_device(c->device));
    FLI_CK(fl_d2d(c->d_value, device_values, sizeof(double) * c->n_sites, c->stream));
    FLI_CK(fl_stream_sync(c->stream));
    c->has_values = true;
    return FASTLEM_OK;
}

int fastlem_interp_set_values_from(fastlem_interp* c, fastlem_ctx* solver) {
    if (!c || !solver) return FASTLEM_E_INVALID;
    if (fastlem_get_device(solver) != c->device)
        return fail(c, FASTLEM_E_INVALID, "set_values_from: the solver context lives on another device");
    FLI_CK(fl_set_device(c->device));
    // the solver writes its elevations (caller's numbering) straight into the interpolator's value array
    int rc = fastlem_download_to_device(solver, c->d_value);
    if (rc) return fail(c, rc, std::string("set_values_from: ") + fastlem_last_error(solver));
    c->has_values = true;
    return FASTLEM_OK;
}

int fastlem_interp_points(fastlem_interp* c, uint32_t n_points, const double* points_xy, double* out) {
    if (!c || (n_points && (!points_xy || !out))) return FASTLEM_E_INVALID;
    if (!c->has_values) return fail(c, FASTLEM_E_STATE, "interpolate: values not set");
    if (n_points == 0) return FASTLEM_OK;
    FLI_CK(fl_set_device(c->device));
    int rc = ensure_out(c, n_points);
    if (rc) return rc;
    if (n_points > c->query_cap) {
        if (c->d_query) fl_free(c->d_query, c->stream);
        c->d_query = nullptr;
        c->query_cap = 0;
        FLI_CK(ialloc(c->d_query, n_points));
        c->query_cap = n_points;
    }
    FLI_CK(fl_h2d(c->d_query, points_xy, sizeof(double) * 2 * (size_t)n_points, c->stream));
    FLI_CK(fl_memset(c->d_flags, 0, sizeof(uint32_t) * FLI_N_FLAGS, c->stream));
    FLI_CK(fl_event_record(c->ev[0], c->stream));
    FLI_LAUNCH_N(k_nn_points, n_points, model_of(c), n_points, c->d_query, c->d_out, c->d_flags);
    FLI_CK(fl_event_record(c->ev[1], c->stream));
    FLI_CK(fl_d2h(out, c->d_out, sizeof(double) * n_points, c->stream));
    return finish_query(c, n_points);
}

int fastlem_interp_raster(fastlem_interp* c, const fastlem_raster* r, double* out) {
    if (!c) return FASTLEM_E_INVALID;
    int rc = validate_raster(c, r);
    if (rc) return rc;
    if (!c->has_values) return fail(c, FASTLEM_E_STATE, "raster: values not set");
    const size_t count = (size_t)(r->row_end - r->row_begin) * r->width;
    if (count == 0) return FASTLEM_OK;
    if (!out) return fail(c, FASTLEM_E_INVALID, "raster: null output");
    FLI_CK(fl_set_device(c->device));
    rc = ensure_out(c, count);
    if (rc) return rc;
    rc = launch_raster(c, r, c->d_out);
    if (rc) return rc;
    FLI_CK(fl_d2h(out, c->d_out, sizeof(double) * count, c->stream));
    return finish_query(c, count);
}

int fastlem_interp_raster_device(fastlem_interp* c, const fastlem_raster* r, double* device_out) {
    if (!c) return FASTLEM_E_INVALID;
    int rc = validate_raster(c, r);
    if (rc) return rc;
    if (!c->has_values) return fail(c, FASTLEM_E_STATE, "raster: values not set");
    const size_t count = (size_t)(r->row_end - r->row_begin) * r->width;
    if (count == 0) return FASTLEM_OK;
    if (!device_out) return fail(c, FASTLEM_E_INVALID, "raster: null output");
    FLI_CK(fl_set_device(c->device));
    rc = launch_raster(c, r, device_out);
    if (rc) return rc;
    return finish_query(c, count);
}

int fastlem_interp_get_stats(const fastlem_interp* c, fastlem_interp_stats* out) {
    if (!c || !out) return FASTLEM_E_INVALID;
    *out = c->stats;
    return FASTLEM_OK;
}

}  // extern "C"
