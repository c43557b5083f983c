/* fastlem_b200.h -- C ABI of the B200-native terrain solve.
 *
 * Drop-in boundary for ONE path of TadaTeruki/fastlem 0.1.4: the body of
 * `TerrainGenerator::generate()` (reference src/lem/generator.rs:90-213, with
 * src/lem/stream_tree.rs:72-243 and src/lem/drainage_basin.rs:13-46 underneath).
 * The reference has no FFI of its own; these are the entry points a Rust shim inside
 * `generate()` binds (see INTEGRATION.md for the `extern "C"` block and the packing code).
 *
 * Conventions
 *   - every function returns FASTLEM_OK (0) or a negative FASTLEM_E_* code; nothing throws
 *     across the boundary; `fastlem_last_error` gives the text of the last failure on a ctx.
 *   - plain pointers and sizes only; all arrays are HOST memory owned by the caller.
 *   - one ctx per generate() call (or reused across calls); a ctx owns its CUDA stream and
 *     device buffers and has no global mutable state, so independent ctxs may be used from
 *     different threads concurrently (TerrainGenerator is Clone and has no interior
 *     mutability, generator.rs:36).  A single ctx must not be used from two threads at once.
 *   - node indices and CSR offsets are uint32 (the reference uses usize): n < 2^32-1, D < 2^32.
 *   - there is NO CPU fallback: without a CUDA device `fastlem_create` fails with
 *     FASTLEM_E_CUDA.
 */
#ifndef FASTLEM_B200_H
#define FASTLEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fastlem_ctx fastlem_ctx;

enum {
    FASTLEM_OK = 0,
    FASTLEM_E_INVALID = -1,  /* bad argument (null pointer, index out of range, ...) */
    FASTLEM_E_STATE = -2,    /* call order: graph / parameters not set (generator.rs:91-116 analogues) */
    FASTLEM_E_CUDA = -3,     /* CUDA runtime failure, or no device */
    FASTLEM_E_NOMEM = -4
};

/* Sentinel for "max_iteration not set" (generator.rs:140: max_iteration.unwrap_or(u32::MAX)). */
#define FASTLEM_UNTIL_STABLE 0xFFFFFFFFu

/* Creates a context on CUDA device `device_ordinal`. */
int fastlem_create(fastlem_ctx** out, int device_ordinal);
void fastlem_destroy(fastlem_ctx* ctx);

/* Device memory is taken from a pool the library keeps per device: the buffers of a destroyed context are reused by
 * the next one (a caller that creates a context per generate(), as the Rust shim does, pays the driver's allocation
 * cost once, not per call).  This returns what the pool holds and no live context uses to the driver.  The environment
 * variable FASTLEM_NO_POOL=1 disables the pool (plain cudaMalloc / cudaFree). */
int fastlem_trim_memory(int device_ordinal);
const char* fastlem_last_error(const fastlem_ctx* ctx);
/* CUDA device ordinal the context lives on (-1 for a null context). */
int fastlem_get_device(const fastlem_ctx* ctx);

/* The model, as `generate()` reads it through trait Model (src/core/traits.rs:13-20):
 *   n        = model.num()                                   (generator.rs:99-105)
 *   row_ptr  = n+1 offsets; row i = graph.neighbors_of(i) IN ITERATION ORDER -- the order is
 *              semantically significant (receiver tie-break stream_tree.rs:122-133, BFS order
 *              drainage_basin.rs:23, lake connection stream_tree.rs:204-236)
 *   col/dist = neighbour index and edge attribute (length) per slot
 *   areas    = model.areas()
 * The graph must be simple and symmetric with equal lengths in both directions (what
 * terrain-graph's add_edge produces).  The arrays are copied to HBM before this returns and the
 * pointers are not kept: the caller may free or reuse them at once.
 */
int fastlem_set_graph(fastlem_ctx* ctx, uint32_t n, const uint32_t* row_ptr, const uint32_t* col,
                      const double* dist, const double* areas);

/* Per-site parameters, flattened from Vec<TopographicalParameters> (src/core/parameters.rs:24-30):
 *   initial_elevation = base_elevation[i] + rng.gen::<f64>() * f64::EPSILON  (generator.rs:134-138;
 *                       the shim keeps calling `rand`, so the noise stream stays the crate's own)
 *   erodibility, uplift_rate
 *   tan_max_slope     = tan(max_slope) per site, NaN where max_slope is None; NULL if None everywhere
 *                       (generator.rs:194 evaluates max_slope.tan(); done once on the host)
 *   outlets           = generator.rs:120-132: indices with is_outlet in ascending order, or
 *                       model.default_outlets() if there is none, in that order.
 */
int fastlem_set_parameters(fastlem_ctx* ctx, const double* initial_elevation, const double* erodibility,
                           const double* uplift_rate, const double* tan_max_slope, const uint32_t* outlets,
                           uint32_t n_outlets);

/* The loop of generator.rs:140-210: runs until no elevation changes or `max_iteration` bodies have
 * run (FASTLEM_UNTIL_STABLE = not set).  Entirely on the device.  Writes the final elevations (n
 * doubles, host) and the number of loop bodies executed.  May be called again: every call restarts
 * from the uploaded initial elevations.
 */
int fastlem_generate(fastlem_ctx* ctx, uint32_t max_iteration, double* elevations_out, uint32_t* iterations_done);

/* fastlem_generate split in two, so a caller can keep results on the device (bench, ensembles):
 * `fastlem_run` leaves the elevations in HBM, `fastlem_download` copies them out. */
int fastlem_run(fastlem_ctx* ctx, uint32_t max_iteration, uint32_t* iterations_done);
int fastlem_download(fastlem_ctx* ctx, double* elevations_out);
/* Same as fastlem_download, but into a caller-owned DEVICE buffer of n doubles on the ctx's device
 * (used to hand an ensemble member's result to an NCCL gather without a host round trip). */
int fastlem_download_to_device(fastlem_ctx* ctx, double* device_elevations_out);

/* Options.  "profile" = 1 (default 0) records CUDA events around every stage and fills the ms_* fields of
 * fastlem_stats; "keep_stages" = 1 (default 0) keeps the pre-lake-removal receivers/labels of the last iteration for
 * fastlem_debug_fetch.  The rest select between implementations that give bit-identical results:
 * "sweep" 0 = one launch per tree level, 1 / 2 = path-decomposed (thread / warp per path), 3 = dataflow sweeps
 * (default); "incremental" (1), "incr_div" (16), "first_flow" (1), "fuse_levels" (1), "flood_device" (1), "outlet_closed_form" (1),
 * "park_after" (8), "key_base", "rebuild_every" (0 = adaptive), "rebuild_growth" / "rebuild_height" (0 = by model size: 4 / 150 percent up to 1M sites,
 * 8 / 250 from 16M sites on), "k1_bulk" (1), "overlap" (1), "outlet_closed_form" (1). */
int fastlem_set_option(fastlem_ctx* ctx, const char* name, int64_t value);

typedef struct fastlem_stats {
    uint32_t iterations;       /* loop bodies executed by the last run */
    uint32_t lake_iterations;  /* of which ran lake removal (stream_tree.rs:91-96) */
    uint32_t depth_first;      /* stream-tree depth (levels) in the first / last iteration */
    uint32_t depth_last;
    uint64_t kernel_launches;  /* kernels of this library launched by the last run */
    double ms_run;             /* device time of the last fastlem_run, CUDA events */
    double ms_upload;          /* host->device copies of set_graph + set_parameters (wall clock) */
    double ms_flood_rank;      /* one-off flood-order computation (wall clock; 0 if never needed) */
    double ms_download;
    /* "profile"=1 only: summed over the iterations of the last run */
    double ms_receivers, ms_labels, ms_lakes, ms_order, ms_area, ms_elevation;
    uint64_t n_receivers, n_labels, n_lakes, n_order, n_area, n_elevation; /* launches per stage */
    /* "sweep" >= 1: path / segment layout of the last iteration */
    uint32_t rebuilds;    /* layout rebuilds (site renumberings) during the last run */
    uint32_t path_levels; /* nesting depth of the path decomposition = rounds per sweep */
    uint32_t paths;       /* number of paths */
    uint32_t incremental_iterations; /* iterations whose drainage areas were updated incrementally */
    uint32_t flood_on_device; /* 1: the flood order was computed on the device, 0: exact host replay (ties) */
    uint32_t outlet_ranks_on_device; /* 1: the outlets' own ranks came from their closed form on the device, 0: host replay
                                      * of that prefix (an outlet without unvisited neighbours early in the order) */
    /* "profile"=2 only (synchronises after every bracketed kernel: a diagnostic mode): device time and launch count
     * of single kernels of the default path, summed over the last run; index = FASTLEM_K_* */
    double ms_kernel[8];
    uint64_t n_kernel[8];
} fastlem_stats;
enum {
    FASTLEM_K_RECEIVERS = 0,      /* k_receivers_bulk (K1; k_receivers_mask with "k1_bulk" = 0) */
    FASTLEM_K_AREA_FLOW = 1,      /* k_area_flow: thread-level climbs of a full K4 pass */
    FASTLEM_K_INCR_START = 2,     /* k_incr_start: thread-level climbs of an incremental K4 pass */
    FASTLEM_K_AREA_FLOW_LONG = 3, /* k_area_flow_long: warp-level climbs of the long segments (K4) */
    FASTLEM_K_ELEV_PLAN = 4,      /* k_elev_plan (K5) */
    FASTLEM_K_ELEV_TOP = 5,       /* k_elev_top (K5: the run queue) */
    FASTLEM_K_ELEV_LOW = 6,       /* k_elev_low (K5: one launch per nesting height below the cut) */
    FASTLEM_K_REBUILD = 7         /* every kernel of a site renumbering */
};
int fastlem_get_stats(const fastlem_ctx* ctx, fastlem_stats* out);

/* Stage dumps of the LAST iteration executed, for parity tests.  `bytes` must equal the size of the
 * stage array (n * 4 for uint32 stages, n * 8 for double stages). */
enum {
    FASTLEM_STAGE_RECEIVERS = 0,         /* u32: stream_tree.next after lake removal */
    FASTLEM_STAGE_RECEIVERS_INITIAL = 1, /* u32: next before lake removal ("keep_stages") */
    FASTLEM_STAGE_LABELS_INITIAL = 2,    /* u32: subroot of find_roots_with_lakes ("keep_stages") */
    FASTLEM_STAGE_LABELS = 3,            /* u32: root of the final forest */
    FASTLEM_STAGE_DEPTH = 4,             /* u32: depth in the final forest, 0xFFFFFFFF if unreached */
    FASTLEM_STAGE_DRAINAGE_AREA = 5,     /* f64 */
    FASTLEM_STAGE_RESPONSE_TIME = 6,     /* f64 */
    FASTLEM_STAGE_ELEVATION = 7,         /* f64 */
    FASTLEM_STAGE_FLOOD_RANK = 8         /* u32: pop order of the lake flood (0xFFFFFFFF = never popped / not computed) */
};
int fastlem_debug_fetch(fastlem_ctx* ctx, int stage, void* out, size_t bytes);

/* Host utility for the C++/Python mirrors of TerrainGenerator (NOT needed by the Rust shim, which keeps
 * using the `rand` crate): generator.rs:134-138, out[i] = base[i] + StdRng::seed_from_u64(0).gen::<f64>() * EPSILON. */
void fastlem_host_initial_elevations(uint32_t n, const double* base_elevation, double* out);

/* Host utility for the mirrors: generator.rs:194 compares the slope with `max_slope.tan()`; Rust's f64::tan is the
 * platform libm's tan.  out[i] = tan(max_slope[i]) with libm (NaN = None stays NaN), so that every mirror hands
 * fastlem_set_parameters bit-identical `tan_max_slope` values (a SIMD tan such as numpy's may differ by an ulp). */
void fastlem_host_tan_max_slope(uint32_t n, const double* max_slope, double* out);

/* Host utility: the graph build of TerrainModel2DBulider::build, builder.rs:252-268, written
 * straight into the boundary format of fastlem_set_graph -- for every triangle (a, b, c) of the builder's
 * triangulation, in order, the half-edges a->b, b->c, c->a with from < to become edges of length
 * Site2D::distance (sites.rs:27-29), appended to both endpoints' rows (terrain-graph's add_edge).  The shim calls it on
 * `voronoi.triangulation().triangles` instead of building the Vec<Vec<..>> graph and re-walking neighbors_of.
 * Two calls: with col = dist = NULL it fills row_ptr (n_sites + 1) and *nnz_out; with buffers of `capacity` >= nnz
 * entries it fills everything.  Pure host code, no device needed. */
int fastlem_host_graph_from_triangles(uint32_t n_sites, const double* sites_xy, uint32_t n_triangles,
                                      const uint32_t* triangles, uint32_t* row_ptr, uint32_t* col, double* dist,
                                      uint64_t capacity, uint64_t* nnz_out);

/* ------------------------------------------------------------------------------------------------
 * Terrain2D::get_elevation  
 *
 * Replaces TerrainInterpolator2D (reference src/models/surface/interpolator.rs:6-28), i.e. the calls
 * `naturalneighbor::Interpolator::new(sites)` (interpolator.rs:11-15, made by
 * TerrainModel2D::create_terrain_from_result, model.rs:62-68) and
 * `Interpolator::interpolate(&elevations, point)` (interpolator.rs:17-27, made once per pixel through
 * Terrain2D::get_elevation, terrain.rs:36-38, by every consumer: examples/landscape_evolution.rs:48-59,
 * examples/terrain_generation_advanced.rs:294-315).  The interpolant is Sibson's natural-neighbour
 * interpolation on the Delaunay triangulation of the sites; a query outside the convex hull is `None`,
 * returned here as NaN.
 * ------------------------------------------------------------------------------------------------ */
typedef struct fastlem_interp fastlem_interp;

/* TerrainInterpolator2D::new (interpolator.rs:11-15).  The Delaunay triangulation is handed over in
 * delaunator's layout -- the shim calls `delaunator::triangulate(&points)` exactly as the crate does:
 *   sites_xy   = 2 * n_sites doubles, x0 y0 x1 y1 ...
 *   triangles  = 3 * n_triangles site indices; half-edge e = 3 t + k runs triangles[e] -> triangles[3 t + (k+1)%3]
 *   halfedges  = 3 * n_triangles: index of the opposite half-edge, 0xFFFFFFFF (usize::MAX truncated) on the hull
 * Either orientation is accepted as long as all triangles agree.  Checked on the device: index ranges, half-edge
 * pairing, orientation, the Delaunay property; a violation returns FASTLEM_E_INVALID (text on stderr).
 * The arrays are copied to HBM; the pointers need not stay valid. */
int fastlem_interp_create(fastlem_interp** out, int device_ordinal, uint32_t n_sites, const double* sites_xy,
                          uint32_t n_triangles, const uint32_t* triangles, const uint32_t* halfedges);
void fastlem_interp_destroy(fastlem_interp* it);
const char* fastlem_interp_last_error(const fastlem_interp* it);

/* The `elevations` slice of interpolate() (interpolator.rs:17): n_sites doubles from the host, from a device
 * buffer on the interpolator's device, or straight from a solver context that has run (same device; no host
 * round trip -- Terrain2D built lazily from the generate() result). */
int fastlem_interp_set_values(fastlem_interp* it, const double* values);
int fastlem_interp_set_values_device(fastlem_interp* it, const double* device_values);
int fastlem_interp_set_values_from(fastlem_interp* it, fastlem_ctx* solver);

/* Terrain2D::get_elevation for a batch of points: out[i] = interpolate(values, (x_i, y_i)), NaN = None. */
int fastlem_interp_points(fastlem_interp* it, uint32_t n_points, const double* points_xy, double* out);

/* The per-pixel loop of the examples as one call.  Pixel (col, row) is queried at
 *     x = span_x * ((col + pixel_offset) / width)  + x0
 *     y = span_y * ((row + pixel_offset) / height) + y0
 * (examples/landscape_evolution.rs:49-50: x0 = 0, span = bound_max, offset 0;
 *  examples/terrain_generation_advanced.rs:296-299: offset 0.5).  Rows [row_begin, row_end) are computed --
 * the unit of partitioning across GPUs -- and written row-major, (row_end - row_begin) * width doubles. */
typedef struct fastlem_raster {
    double x0, y0, span_x, span_y, pixel_offset;
    uint32_t width, height, row_begin, row_end;
} fastlem_raster;
int fastlem_interp_raster(fastlem_interp* it, const fastlem_raster* raster, double* out);
/* Same, into a caller-owned DEVICE buffer on the interpolator's device (for an NCCL gather of row blocks). */
int fastlem_interp_raster_device(fastlem_interp* it, const fastlem_raster* raster, double* device_out);

typedef struct fastlem_interp_stats {
    double ms_setup;          /* fastlem_interp_create: uploads, circumcircles, checks, hint grid (wall clock) */
    double ms_query_kernel;   /* device time of the last points / raster kernel, CUDA events */
    uint64_t queries;         /* points or pixels of the last query call */
    uint64_t kernel_launches; /* kernels of this library launched on this interpolator so far */
    uint32_t grid_x, grid_y;  /* hint grid */
    uint32_t grid_passes;     /* dilation passes needed to fill empty hint cells */
    uint32_t clockwise;       /* 1 if the triangles were clockwise */
} fastlem_interp_stats;
int fastlem_interp_get_stats(const fastlem_interp* it, fastlem_interp_stats* out);

/* Library identification: "fastlem_b200 <version> sm_100a" (or "... emu" for the test-only host build). */
const char* fastlem_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FASTLEM_B200_H */
