// fl_interp.cuh -- natural-neighbour (Sibson) interpolation of site elevations: `Terrain2D::get_elevation`
// (reference src/models/surface/terrain.rs:36-38 -> src/models/surface/interpolator.rs:17-27 ->
// naturalneighbor::Interpolator::interpolate; the crate `naturalneighbor` 1.2.2 is an un-vendored
// dependency, so what is restated here is the published definition of the interpolant:
//
//     z(p) = sum_i w_i z_i,   w_i = area(V'(p) & V(i)) / area(V'(p))
//
// V(i) the Voronoi cell of site i, V'(p) the cell p would get if it were inserted; `None` outside the
// convex hull of the sites).
//
// One thread per query point.
//   1. locate: visibility walk through the Delaunay triangulation from a hint triangle taken from a uniform
//      grid over the bounding box of the sites;
//   2. cavity: the triangles whose circumcircle contains p form a triangulated polygon without interior
//      vertices, so their dual graph is a tree: a depth-first walk that never steps back over the edge it
//      came through needs no visited set;
//   3. weights: the region p steals from site v is the polygon  g_in, C_t1 .. C_tk, g_out  (C_t = circumcentre
//      of the cavity triangles around v, g = circumcentre of (p, cavity-boundary edge at v)).  Its shoelace
//      sum taken about m = (p+v)/2 -- a point ON the closing side g_out -> g_in, which therefore drops out --
//      splits into one term per (cavity triangle, edge), so numerator and denominator are accumulated on
//      the fly: no per-vertex table, no ordering of the boundary.  Only circumcentres of cavity triangles and
//      of (p, boundary edge) are used; p is never collinear with a boundary edge of the cavity except on the
//      hull itself (handled: linear interpolation along the hull edge), unlike Watson's per-triangle form.
// Compiled with -fmad=false like the rest of the library; the FL_EMU host build runs the same bodies.
#pragma once
#include "fl_rt.h"

#define FLI_NONE 0xFFFFFFFFu
// Build-time variants of the query kernel, kept for A/B runs on the device (tools/ab_raster.py).  The defaults are the
// fastest combination measured on a B200 (profiles/r1c_ab_raster*.txt: 4096^2 pixels over 1M sites, 2.83 ms with
// everything off -> 2.02 ms):
//   FLI_MINBLOCKS n : __launch_bounds__(256, n) on the query kernels (0 = ptxas' choice: 80 registers, 3 CTAs per SM;
//                     4 = 64 registers with ~100 B of spills, +12 %; 5 and 6 lose to their spills)
//   FLI_RECIP 1     : one reciprocal per (p, boundary edge) circumcentre instead of two divisions (+12 %)
//   FLI_UNIFIED 1   : interior and boundary edges share the code of the first term (+6 % on top of FLI_RECIP)
//   FLI_WARP_8X4 1  : raster lanes cover 8 x 4 pixels instead of 16 x 2 (+2 %)
#ifndef FLI_MINBLOCKS
#define FLI_MINBLOCKS 4
#endif
#ifndef FLI_RECIP
#define FLI_RECIP 1
#endif
#ifndef FLI_UNIFIED
#define FLI_UNIFIED 1
#endif
#ifndef FLI_WARP_8X4
#define FLI_WARP_8X4 1
#endif
//   FLI_DENORM 1    : vertex coordinates stored per triangle (48 B more per triangle): the loads of a triangle's record
//                     no longer wait for its vertex ids (+3 %)
#ifndef FLI_DENORM
#define FLI_DENORM 1
#endif
#if FLI_MINBLOCKS > 0 && !defined(FL_EMU)
#define FLI_QUERY_BOUNDS __launch_bounds__(256, FLI_MINBLOCKS)
#else
#define FLI_QUERY_BOUNDS __launch_bounds__(256)
#endif
#ifndef FLI_STACK
#define FLI_STACK 48      // depth-first stack of the cavity walk (pending triangles)
#endif
#define FLI_MAX_CAVITY 512  // triangles visited per query before the walk is declared broken

// flag words
enum { FLI_F_BAD_INDEX = 0, FLI_F_POS = 1, FLI_F_NEG = 2, FLI_F_DEGENERATE = 3, FLI_F_NOT_DELAUNAY = 4,
       FLI_F_EMPTY_CELLS = 5, FLI_F_WALK_OVERFLOW = 6, FLI_F_CAVITY_OVERFLOW = 7, FLI_F_BAD_HALFEDGE = 8,
       FLI_N_FLAGS = 16 };

struct alignas(16) FliPt { double x, y; };
struct alignas(16) FliTri { uint32_t v[3]; uint32_t pad; };   // vertices; edge k runs v[k] -> v[(k+1)%3]
struct alignas(16) FliNbr { uint32_t t[3]; uint32_t pad; };   // triangle across edge k, FLI_NONE on the hull
struct alignas(16) FliCirc { double x, y, r2, pad; };         // circumcentre and squared circumradius
struct alignas(16) FliGeo { FliPt p[3]; };                    // FLI_DENORM: the three vertices' coordinates

struct FliGrid {
    double x0, y0, inv_cell_x, inv_cell_y;
    uint32_t gx, gy;
};

#define FLI_TID (blockIdx.x * blockDim.x + threadIdx.x)

#ifdef FL_EMU
inline double fli_nan() { return std::nan(""); }
#else
__device__ __forceinline__ double fli_nan() { return __longlong_as_double(0x7FF8000000000000ll); }
#endif

__device__ __forceinline__ double fli_cross(double ax, double ay, double bx, double by) { return ax * by - ay * bx; }

// circumcentre of (a, b, c) relative to a; false when the three points are collinear
__device__ __forceinline__ bool fli_circumcentre(double ax, double ay, double bx, double by, double cx, double cy,
                                                 double* ux, double* uy) {
    const double ex = bx - ax, ey = by - ay, fx = cx - ax, fy = cy - ay;
    const double d = 2.0 * (ex * fy - ey * fx);
    const double e2 = ex * ex + ey * ey, f2 = fx * fx + fy * fy;
    *ux = (fy * e2 - ey * f2) / d;
    *uy = (ex * f2 - fx * e2) / d;
    return d != 0.0;
}

// circumcentre of (p, a, b) relative to p for the query kernels
__device__ __forceinline__ bool fli_circumcentre_q(double ax, double ay, double bx, double by, double cx, double cy,
                                                   double* ux, double* uy) {
#if FLI_RECIP
    const double ex = bx - ax, ey = by - ay, fx = cx - ax, fy = cy - ay;
    const double d = 2.0 * (ex * fy - ey * fx);
    const double e2 = ex * ex + ey * ey, f2 = fx * fx + fy * fy;
    const double inv = 1.0 / d;
    *ux = (fy * e2 - ey * f2) * inv;
    *uy = (ex * f2 - fx * e2) * inv;
    return d != 0.0;
#else
    return fli_circumcentre(ax, ay, bx, by, cx, cy, ux, uy);
#endif
}

// ------------------------------------------------------------------------------------------------
// set-up kernels (once per triangulation)
// ------------------------------------------------------------------------------------------------
// sites relative to the lower corner of their bounding box (exact when the offset is large against the extent)
__global__ void __launch_bounds__(256) k_nn_translate(uint32_t n, FliPt* __restrict__ site, double ox, double oy) {
    const uint32_t i = FLI_TID;
    if (i >= n) return;
    FliPt p = site[i];
    p.x -= ox;
    p.y -= oy;
    site[i] = p;
}

// delaunator's arrays (triangles[3T], halfedges[3T]; usize::MAX = no opposite half-edge) -> FliTri / FliNbr,
// circumcircles, orientation census.
__global__ void __launch_bounds__(256) k_nn_prepare(uint32_t n_sites, uint32_t n_tri, const FliPt* __restrict__ site,
                                                    const uint32_t* __restrict__ triangles,
                                                    const uint32_t* __restrict__ halfedges, FliTri* __restrict__ tri,
                                                    FliNbr* __restrict__ nbr, FliCirc* __restrict__ circ,
                                                    FliGeo* __restrict__ geo, uint32_t* __restrict__ flags) {
    const uint32_t t = FLI_TID;
    if (t >= n_tri) return;
    FliTri T;
    FliNbr N;
    bool ok = true;
    for (int k = 0; k < 3; ++k) {
        T.v[k] = triangles[3u * t + k];
        ok = ok && T.v[k] < n_sites;
        const uint32_t h = halfedges[3u * t + k];
        if (h == FLI_NONE) {
            N.t[k] = FLI_NONE;
        } else if (h >= 3u * n_tri) {
            N.t[k] = FLI_NONE;
            atomicOr(&flags[FLI_F_BAD_HALFEDGE], 1u);
        } else {
            N.t[k] = h / 3u;
            // the opposite half-edge must run the other way between the same two sites
            const uint32_t k2 = h % 3u, t2 = h / 3u;
            const uint32_t a2 = triangles[3u * t2 + k2], b2 = triangles[3u * t2 + (k2 + 1u) % 3u];
            if (a2 != triangles[3u * t + (k + 1) % 3] || b2 != triangles[3u * t + k] || t2 == t)
                atomicOr(&flags[FLI_F_BAD_HALFEDGE], 1u);
        }
    }
    T.pad = 0;
    N.pad = 0;
    tri[t] = T;
    nbr[t] = N;
    FliCirc C;
    C.x = C.y = C.r2 = C.pad = 0.0;
    if (!ok) {
        atomicOr(&flags[FLI_F_BAD_INDEX], 1u);
        circ[t] = C;
        return;
    }
    const FliPt a = site[T.v[0]], b = site[T.v[1]], c = site[T.v[2]];
    if (geo) {
        FliGeo G;
        G.p[0] = a; G.p[1] = b; G.p[2] = c;
        geo[t] = G;
    }
    const double o = fli_cross(b.x - a.x, b.y - a.y, c.x - a.x, c.y - a.y);
    if (o > 0.0) atomicOr(&flags[FLI_F_POS], 1u);
    else if (o < 0.0) atomicOr(&flags[FLI_F_NEG], 1u);
    double ux, uy;
    if (!fli_circumcentre(a.x, a.y, b.x, b.y, c.x, c.y, &ux, &uy) || !(o == o)) {
        atomicOr(&flags[FLI_F_DEGENERATE], 1u);
        circ[t] = C;
        return;
    }
    C.x = a.x + ux;
    C.y = a.y + uy;
    C.r2 = ux * ux + uy * uy;
    circ[t] = C;
}

// Delaunay check: the vertex opposite to every interior edge must not lie strictly inside the circumcircle.
__global__ void __launch_bounds__(256) k_nn_check_delaunay(uint32_t n_tri, const FliPt* __restrict__ site,
                                                           const FliTri* __restrict__ tri, const FliNbr* __restrict__ nbr,
                                                           const FliCirc* __restrict__ circ, uint32_t* __restrict__ flags) {
    const uint32_t t = FLI_TID;
    if (t >= n_tri) return;
    const FliTri T = tri[t];
    const FliNbr N = nbr[t];
    const FliCirc C = circ[t];
    for (int k = 0; k < 3; ++k) {
        const uint32_t t2 = N.t[k];
        if (t2 == FLI_NONE) continue;
        const FliTri T2 = tri[t2];
        const uint32_t a = T.v[k], b = T.v[(k + 1) % 3];
        for (int j = 0; j < 3; ++j) {
            const uint32_t w = T2.v[j];
            if (w == a || w == b) continue;
            const FliPt q = site[w];
            const double dx = q.x - C.x, dy = q.y - C.y;
            // tolerance: circumcircles of hull slivers are computed with a relative error that grows with their
            // radius; the check is there to refuse grossly non-Delaunay input, not to certify exactness
            if (dx * dx + dy * dy < C.r2 * (1.0 - 1e-6)) atomicAdd(&flags[FLI_F_NOT_DELAUNAY], 1u);
        }
    }
}

__device__ __forceinline__ uint32_t fli_cell(const FliGrid g, double x, double y) {
    double fx = (x - g.x0) * g.inv_cell_x, fy = (y - g.y0) * g.inv_cell_y;
    // clamp in floating point first: queries may be far outside the box (or NaN -> cell 0)
    fx = fx > 0.0 ? fx : 0.0;
    fy = fy > 0.0 ? fy : 0.0;
    uint32_t cx = fx < (double)g.gx ? (uint32_t)fx : g.gx - 1u;
    uint32_t cy = fy < (double)g.gy ? (uint32_t)fy : g.gy - 1u;
    return cy * g.gx + cx;
}

// hint grid: every cell takes the smallest triangle id whose centroid falls into it
__global__ void __launch_bounds__(256) k_nn_grid_fill(uint32_t n_tri, const FliPt* __restrict__ site,
                                                      const FliTri* __restrict__ tri, FliGrid g,
                                                      uint32_t* __restrict__ cell) {
    const uint32_t t = FLI_TID;
    if (t >= n_tri) return;
    const FliTri T = tri[t];
    const FliPt a = site[T.v[0]], b = site[T.v[1]], c = site[T.v[2]];
    atomicMin(&cell[fli_cell(g, (a.x + b.x + c.x) / 3.0, (a.y + b.y + c.y) / 3.0)], t);
}

// one dilation pass over the empty cells (double buffered, smallest id of the 8 neighbours: deterministic)
__global__ void __launch_bounds__(256) k_nn_grid_dilate(FliGrid g, const uint32_t* __restrict__ src,
                                                        uint32_t* __restrict__ dst, uint32_t* __restrict__ flags) {
    const uint32_t i = FLI_TID;
    if (i >= g.gx * g.gy) return;
    uint32_t v = src[i];
    if (v == FLI_NONE) {
        const int cx = (int)(i % g.gx), cy = (int)(i / g.gx);
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int x = cx + dx, y = cy + dy;
                if (x < 0 || y < 0 || x >= (int)g.gx || y >= (int)g.gy) continue;
                const uint32_t w = src[(uint32_t)y * g.gx + (uint32_t)x];
                v = w < v ? w : v;
            }
        if (v == FLI_NONE) atomicAdd(&flags[FLI_F_EMPTY_CELLS], 1u);
    }
    dst[i] = v;
}

// ------------------------------------------------------------------------------------------------
// the query
// ------------------------------------------------------------------------------------------------
struct FliModel {
    const FliPt* site;
    const FliTri* tri;
    const FliNbr* nbr;
    const FliCirc* circ;
    const FliGeo* geo;  // FLI_DENORM only
    const uint32_t* cell;
    const double* value;
    FliGrid grid;
    uint32_t n_tri;
    uint32_t max_walk;
    double sgn;  // +1: triangles counter-clockwise, -1: clockwise
    double ox, oy;  // origin of the device-side coordinates (lower corner of the sites' bounding box)
};

// Returns the interpolated value, or NaN for `None` (outside the convex hull; NaN coordinates).
__device__ __forceinline__ double fli_query(const FliModel& M, const double qx, const double qy,
                                            uint32_t* __restrict__ flags) {
    const double nan = fli_nan();
    if (!(qx == qx) || !(qy == qy) || M.n_tri == 0u) return nan;
    const double px = qx - M.ox, py = qy - M.oy;  // same translation as the sites (k_nn_translate)
    // ---- 1. locate -------------------------------------------------------------------------------
    uint32_t t = M.cell[fli_cell(M.grid, px, py)];
    FliTri T;
    FliNbr N;
    FliPt P[3];
    uint32_t steps = 0, came_from = FLI_NONE;
    for (;;) {
        T = M.tri[t];
        N = M.nbr[t];
#if FLI_DENORM
        { const FliGeo G = M.geo[t]; P[0] = G.p[0]; P[1] = G.p[1]; P[2] = G.p[2]; }
#else
        P[0] = M.site[T.v[0]];
        P[1] = M.site[T.v[1]];
        P[2] = M.site[T.v[2]];
#endif
        // The walk never steps straight back over the edge it came through: the two triangles evaluate their shared
        // edge from different base vertices, so for a query on that edge within rounding BOTH can see a tiny negative
        // orientation and the walk would bounce between them until max_walk.  In exact arithmetic a step back is
        // impossible (the edge was crossed because p lies strictly beyond it), so that edge is not a candidate.
        double worst = 0.0;
        int kw = -1;
        for (int k = 0; k < 3; ++k) {
            const FliPt a = P[k], b = P[(k + 1) % 3];
            const double o = M.sgn * fli_cross(b.x - a.x, b.y - a.y, px - a.x, py - a.y);
            if (o < worst && !(came_from != FLI_NONE && N.t[k] == came_from)) { worst = o; kw = k; }
        }
        if (kw < 0) break;  // inside or on the boundary of t
        const uint32_t t2 = N.t[kw];
        if (t2 == FLI_NONE) return nan;  // beyond a hull edge: outside the convex hull
        came_from = t;
        t = t2;
        if (++steps > M.max_walk) {
            atomicOr(&flags[FLI_F_WALK_OVERFLOW], 1u);
            return nan;
        }
    }
    // p on a site: the interpolant is the site's value (weights degenerate to 1)
    for (int k = 0; k < 3; ++k)
        if (P[k].x == px && P[k].y == py) return M.value[T.v[k]];

    // ---- 2 + 3. cavity walk with on-the-fly accumulation ------------------------------------------
    uint32_t st_t[FLI_STACK], st_from[FLI_STACK];
    int sp = 0;
    uint32_t from = FLI_NONE, visited = 0;
    double num = 0.0, den = 0.0;
    for (;;) {
        const FliCirc C = M.circ[t];
        for (int k = 0; k < 3; ++k) {
            const int k1 = (k + 1) % 3;
            const FliPt a = P[k], b = P[k1];
            const uint32_t t2 = N.t[k];
            bool inside = false;
            FliCirc C2;
            C2.x = C2.y = 0.0;
            if (t2 != FLI_NONE) {
                C2 = M.circ[t2];
                if (t2 == from) {
                    inside = true;
                } else {
                    const double dx = px - C2.x, dy = py - C2.y;
                    inside = dx * dx + dy * dy < C2.r2;
                    if (inside) {
                        if (sp < FLI_STACK) { st_t[sp] = t2; st_from[sp] = t; ++sp; }
                        else atomicOr(&flags[FLI_F_CAVITY_OVERFLOW], 1u);
                    }
                }
            }
            const double max_ = 0.5 * (px + a.x), may = 0.5 * (py + a.y);
#if FLI_UNIFIED
            // X = the other end of the Voronoi edge piece that starts at C: the neighbour's circumcentre (interior edge)
            // or g = circumcentre of (p, a, b) (boundary edge of the cavity)
            double Xx = C2.x, Xy = C2.y;
            if (!inside) {
                double ux, uy;
                if (!fli_circumcentre_q(px, py, a.x, a.y, b.x, b.y, &ux, &uy)) {
                    const double ex = b.x - a.x, ey = b.y - a.y;
                    const double s = ((px - a.x) * ex + (py - a.y) * ey) / (ex * ex + ey * ey);
                    return M.value[T.v[k]] + s * (M.value[T.v[k1]] - M.value[T.v[k]]);
                }
                Xx = px + ux;
                Xy = py + uy;
            }
            const double ta = fli_cross(Xx - max_, Xy - may, C.x - max_, C.y - may);
            num += ta * M.value[T.v[k]];
            den += ta;
            if (!inside) {
                const double mbx = 0.5 * (px + b.x), mby = 0.5 * (py + b.y);
                const double tb = fli_cross(C.x - mbx, C.y - mby, Xx - mbx, Xy - mby);
                num += tb * M.value[T.v[k1]];
                den += tb;
            }
#else
            if (inside) {
                // interior edge a -> b (t on its left, t2 on its right): around a, t2 precedes t
                const double term = fli_cross(C2.x - max_, C2.y - may, C.x - max_, C.y - may);
                num += term * M.value[T.v[k]];
                den += term;
            } else {
                // boundary edge of the cavity: t is the first triangle of a's fan and the last of b's
                double ux, uy;
                if (!fli_circumcentre_q(px, py, a.x, a.y, b.x, b.y, &ux, &uy)) {
                    // p on the line through a, b: only possible on a hull edge -> linear along the edge
                    const double ex = b.x - a.x, ey = b.y - a.y;
                    const double s = ((px - a.x) * ex + (py - a.y) * ey) / (ex * ex + ey * ey);
                    return M.value[T.v[k]] + s * (M.value[T.v[k1]] - M.value[T.v[k]]);
                }
                const double gx = px + ux, gy = py + uy;
                const double mbx = 0.5 * (px + b.x), mby = 0.5 * (py + b.y);
                const double ta = fli_cross(gx - max_, gy - may, C.x - max_, C.y - may);
                const double tb = fli_cross(C.x - mbx, C.y - mby, gx - mbx, gy - mby);
                num += ta * M.value[T.v[k]];
                den += ta;
                num += tb * M.value[T.v[k1]];
                den += tb;
            }
#endif
        }
        if (sp == 0) break;
        if (++visited > FLI_MAX_CAVITY) {
            atomicOr(&flags[FLI_F_CAVITY_OVERFLOW], 1u);
            return nan;
        }
        --sp;
        t = st_t[sp];
        from = st_from[sp];
        T = M.tri[t];
        N = M.nbr[t];
#if FLI_DENORM
        { const FliGeo G = M.geo[t]; P[0] = G.p[0]; P[1] = G.p[1]; P[2] = G.p[2]; }
#else
        P[0] = M.site[T.v[0]];
        P[1] = M.site[T.v[1]];
        P[2] = M.site[T.v[2]];
#endif
    }
    return num / den;
}

// arbitrary points (Terrain2D::get_elevation, one call per point in the reference)
__global__ void FLI_QUERY_BOUNDS k_nn_points(FliModel M, uint32_t nq, const FliPt* __restrict__ q,
                                                   double* __restrict__ out, uint32_t* __restrict__ flags) {
    const uint32_t i = FLI_TID;
    if (i >= nq) return;
    const FliPt p = q[i];
    out[i] = fli_query(M, p.x, p.y, flags);
}

// raster: pixel (col, row) -> x = span_x * ((col + offset) / width) + x0, y likewise (the examples' formula:
// examples/landscape_evolution.rs:49-50 with offset 0, examples/terrain_generation_advanced.rs:296-299 with 0.5).
// One 16 x 16 pixel tile per CTA (a warp covers 16 x 2 pixels, so its lanes walk the same few triangles).
struct FliRaster {
    double x0, y0, span_x, span_y, offset;
    uint32_t width, height, row_begin, row_end;
};

__global__ void FLI_QUERY_BOUNDS k_nn_raster(FliModel M, FliRaster R, double* __restrict__ out,
                                                   uint32_t* __restrict__ flags) {
    const uint32_t tiles_x = (R.width + 15u) / 16u;
    const uint32_t tile = blockIdx.x;
#if FLI_WARP_8X4
    // lanes 0..31 of a warp: 8 columns x 4 rows; warps 0..7 of the CTA: 2 x 4 such blocks
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t lx = (warp & 1u) * 8u + (lane & 7u), ly = (warp >> 1) * 4u + (lane >> 3);
#else
    const uint32_t lx = threadIdx.x & 15u, ly = threadIdx.x >> 4;
#endif
    const uint32_t col = (tile % tiles_x) * 16u + lx;
    const uint32_t row = R.row_begin + (tile / tiles_x) * 16u + ly;
    if (col >= R.width || row >= R.row_end) return;
    const double x = R.span_x * (((double)col + R.offset) / (double)R.width) + R.x0;
    const double y = R.span_y * (((double)row + R.offset) / (double)R.height) + R.y0;
    out[(size_t)(row - R.row_begin) * R.width + col] = fli_query(M, x, y, flags);
}
