// nn_oracle.cpp -- CPU oracle for Terrain2D::get_elevation.  TEST INFRASTRUCTURE ONLY: loaded by tests/,
// __graft_entry__.smoke() and bench.py's CPU legs through oracle/oracle.py; the product never links it.
//
// Reference call chain: src/models/surface/terrain.rs:36-38 (get_elevation) -> src/models/surface/interpolator.rs:17-27
// (TerrainInterpolator2D::interpolate) -> naturalneighbor::Interpolator::interpolate (crate `naturalneighbor`,
// caret requirement 1.2.2 in Cargo.toml:15; NOT vendored under /root/reference, no Cargo.lock).
//
// PARITY UNPINNED: the dependency's source is absent and the reference's tests hold no numeric expectation for this
// call (tests/*.rs only render image.png), so what is restated is the published algorithm the crate documents --
// Sibson's natural-neighbour interpolation over the Delaunay triangulation of the sites:
//     1. find the triangle containing the query p; outside the convex hull -> None (NaN here);
//     2. Bowyer-Watson cavity: the connected set of triangles whose circumcircle contains p;
//     3. the boundary of the cavity lists the natural neighbours v_0..v_{k-1}; the area p steals from the Voronoi
//        cell of v_i is the polygon  cc(p, v_i, v_{i+1}), cc(t) for the cavity triangles t around v_i in order,
//        cc(p, v_{i-1}, v_i);   w_i = area_i / sum_j area_j,  z(p) = sum_i w_i z_i.
// Deliberately written differently from the device code (fastlem_b200/csrc/fl_interp.cuh): brute-force point
// location, visited-set BFS with the determinant in-circle predicate, explicit ordered polygons with the shoelace
// sum about v_i, adjacency rebuilt from the triangles alone, long double arithmetic.
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace {

typedef long double real;

struct Pt { real x, y; };

inline real cross(Pt a, Pt b, Pt c) { return (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x); }

inline bool circumcentre(Pt a, Pt b, Pt c, Pt* out) {
    const real ex = b.x - a.x, ey = b.y - a.y, fx = c.x - a.x, fy = c.y - a.y;
    const real d = 2 * (ex * fy - ey * fx);
    if (d == 0) return false;
    const real e2 = ex * ex + ey * ey, f2 = fx * fx + fy * fy;
    out->x = a.x + (fy * e2 - ey * f2) / d;
    out->y = a.y + (ex * f2 - fx * e2) / d;
    return true;
}

// > 0 iff p lies strictly inside the circumcircle of the counter-clockwise triangle (a, b, c)
inline real incircle(Pt a, Pt b, Pt c, Pt p) {
    const real ax = a.x - p.x, ay = a.y - p.y, bx = b.x - p.x, by = b.y - p.y, cx = c.x - p.x, cy = c.y - p.y;
    const real a2 = ax * ax + ay * ay, b2 = bx * bx + by * by, c2 = cx * cx + cy * cy;
    return ax * (by * c2 - b2 * cy) - ay * (bx * c2 - b2 * cx) + a2 * (bx * cy - by * cx);
}

struct Mesh {
    uint32_t n = 0, nt = 0;
    Pt origin{0, 0};  // first site: all coordinates are taken relative to it (exact in long double)
    std::vector<Pt> pts;
    std::vector<uint32_t> tri;                      // counter-clockwise after normalisation
    std::unordered_map<uint64_t, uint32_t> edge_tri;  // directed edge (a -> b) -> triangle having it
    std::vector<uint32_t> nbr;                      // nbr[3t+k] = triangle across edge k of t (from edge_tri)
    uint64_t key(uint32_t a, uint32_t b) const { return ((uint64_t)a << 32) | b; }
    uint32_t tri_of(uint32_t a, uint32_t b) const {
        auto it = edge_tri.find(key(a, b));
        return it == edge_tri.end() ? 0xFFFFFFFFu : it->second;
    }
};

Mesh build_mesh(uint32_t n, const double* xy, uint32_t nt, const uint32_t* tri) {
    Mesh m;
    m.n = n;
    m.nt = nt;
    m.pts.resize(n);
    if (n) m.origin = Pt{(real)xy[0], (real)xy[1]};
    for (uint32_t i = 0; i < n; ++i) m.pts[i] = Pt{(real)xy[2 * i] - m.origin.x, (real)xy[2 * i + 1] - m.origin.y};
    m.tri.assign(tri, tri + 3 * (size_t)nt);
    for (uint32_t t = 0; t < nt; ++t) {
        uint32_t* v = &m.tri[3 * (size_t)t];
        if (cross(m.pts[v[0]], m.pts[v[1]], m.pts[v[2]]) < 0) { uint32_t s = v[1]; v[1] = v[2]; v[2] = s; }
        for (int k = 0; k < 3; ++k) m.edge_tri[m.key(v[k], v[(k + 1) % 3])] = t;
    }
    m.nbr.resize(3 * (size_t)nt);
    for (uint32_t t = 0; t < nt; ++t)
        for (int k = 0; k < 3; ++k) m.nbr[3 * (size_t)t + k] = m.tri_of(m.tri[3 * (size_t)t + (k + 1) % 3], m.tri[3 * (size_t)t + k]);
    return m;
}

double g_mesh_seconds = 0.0;

// Point location for the timed CPU baseline (bench.py): visibility walk from the previous query's triangle, the way a
// CPU implementation scanning an image would do it.  Returns NONE when the walk leaves the hull or does not settle.
uint32_t walk_locate(const Mesh& m, Pt p, uint32_t start) {
    uint32_t t = start < m.nt ? start : 0;
    for (uint32_t steps = 0; steps < 4 * m.nt + 16; ++steps) {
        const uint32_t* v = &m.tri[3 * (size_t)t];
        int worst = -1;
        real wv = 0;
        for (int k = 0; k < 3; ++k) {
            const real o = cross(m.pts[v[k]], m.pts[v[(k + 1) % 3]], p);
            if (o < wv) { wv = o; worst = k; }
        }
        if (worst < 0) return t;
        const uint32_t t2 = m.nbr[3 * (size_t)t + worst];
        if (t2 == 0xFFFFFFFFu) return 0xFFFFFFFFu;
        t = t2;
    }
    return 0xFFFFFFFFu;
}

// weights of one query; returns false for None.  ids/w receive the natural neighbours and their weights.
// `hint`: null = brute-force location; else walk from *hint and store the triangle found.
bool query(const Mesh& m, Pt p, std::vector<uint32_t>* ids, std::vector<real>* w, uint32_t* hint = nullptr) {
    ids->clear();
    w->clear();
    if (!(p.x == p.x) || !(p.y == p.y)) return false;
    // 1. brute-force location: first triangle (in index order) with p inside or on its boundary
    uint32_t t0 = 0xFFFFFFFFu;
    if (hint) {
        t0 = walk_locate(m, p, *hint);
        if (t0 != 0xFFFFFFFFu) *hint = t0;
    }
    for (uint32_t t = 0; !hint && t < m.nt && t0 == 0xFFFFFFFFu; ++t) {
        const uint32_t* v = &m.tri[3 * (size_t)t];
        if (cross(m.pts[v[0]], m.pts[v[1]], p) >= 0 && cross(m.pts[v[1]], m.pts[v[2]], p) >= 0 &&
            cross(m.pts[v[2]], m.pts[v[0]], p) >= 0)
            t0 = t;
    }
    if (t0 == 0xFFFFFFFFu) return false;
    for (int k = 0; k < 3; ++k) {
        const uint32_t v = m.tri[3 * (size_t)t0 + k];
        if (m.pts[v].x == p.x && m.pts[v].y == p.y) { ids->push_back(v); w->push_back(1); return true; }
    }
    // 2. cavity by breadth-first search with a visited set
    std::vector<uint32_t> cav{t0};
    std::unordered_map<uint32_t, bool> in_cav{{t0, true}};
    for (size_t h = 0; h < cav.size(); ++h) {
        for (int k = 0; k < 3; ++k) {
            const uint32_t t2 = m.nbr[3 * (size_t)cav[h] + k];
            if (t2 == 0xFFFFFFFFu || in_cav.count(t2)) continue;
            const uint32_t* u = &m.tri[3 * (size_t)t2];
            const bool in = incircle(m.pts[u[0]], m.pts[u[1]], m.pts[u[2]], p) > 0;
            in_cav[t2] = in;
            if (in) cav.push_back(t2);
        }
    }
    auto inside = [&](uint32_t t) { auto it = in_cav.find(t); return it != in_cav.end() && it->second; };
    // 3. boundary cycle of the cavity: directed edges whose twin's triangle is not in the cavity
    std::unordered_map<uint32_t, uint32_t> next_of, prev_of;
    for (uint32_t t : cav) {
        const uint32_t* v = &m.tri[3 * (size_t)t];
        for (int k = 0; k < 3; ++k) {
            const uint32_t a = v[k], b = v[(k + 1) % 3];
            const uint32_t t2 = m.tri_of(b, a);
            if (t2 != 0xFFFFFFFFu && inside(t2)) continue;
            if (t2 == 0xFFFFFFFFu && cross(m.pts[a], m.pts[b], p) == 0) {
                // p on a hull edge: the limit of the weights is linear interpolation along the edge
                const real ex = m.pts[b].x - m.pts[a].x, ey = m.pts[b].y - m.pts[a].y;
                const real s = ((p.x - m.pts[a].x) * ex + (p.y - m.pts[a].y) * ey) / (ex * ex + ey * ey);
                ids->assign({a, b});
                w->assign({1 - s, s});
                return true;
            }
            next_of[a] = b;
            prev_of[b] = a;
        }
    }
    real total = 0;
    for (auto& kv : next_of) {
        const uint32_t v = kv.first, vnext = kv.second, vprev = prev_of[v];
        const Pt o = m.pts[v];
        std::vector<Pt> poly;
        Pt g;
        if (!circumcentre(p, o, m.pts[vnext], &g)) return false;
        poly.push_back(g);
        // fan of cavity triangles around v, counter-clockwise, starting at the one holding edge v -> vnext
        uint32_t b = vnext;
        for (;;) {
            const uint32_t t = m.tri_of(v, b);
            if (t == 0xFFFFFFFFu || !inside(t)) return false;  // broken cavity
            const uint32_t* u = &m.tri[3 * (size_t)t];
            Pt c;
            if (!circumcentre(m.pts[u[0]], m.pts[u[1]], m.pts[u[2]], &c)) return false;
            poly.push_back(c);
            uint32_t third = 0;
            for (int k = 0; k < 3; ++k)
                if (u[k] == v) third = u[(k + 2) % 3];  // (v, b, third) counter-clockwise
            if (third == vprev) break;
            b = third;
            if (poly.size() > 4096) return false;
        }
        if (!circumcentre(p, m.pts[vprev], o, &g)) return false;
        poly.push_back(g);
        real a2 = 0;  // shoelace about v
        for (size_t i = 0; i < poly.size(); ++i) {
            const Pt q = poly[i], r = poly[(i + 1) % poly.size()];
            a2 += (q.x - o.x) * (r.y - o.y) - (q.y - o.y) * (r.x - o.x);
        }
        ids->push_back(v);
        w->push_back(a2);
        total += a2;
    }
    for (real& x : *w) x /= total;
    return true;
}

}  // namespace

extern "C" {

// out[i] = interpolated value at (qxy[2i], qxy[2i+1]) or NaN for None.
void fo_nn_interpolate(uint32_t n, const double* xy, uint32_t nt, const uint32_t* tri, const double* values,
                       uint32_t nq, const double* qxy, double* out) {
    const Mesh m = build_mesh(n, xy, nt, tri);
    std::vector<uint32_t> ids;
    std::vector<real> w;
    for (uint32_t i = 0; i < nq; ++i) {
        if (!query(m, Pt{(real)qxy[2 * i] - m.origin.x, (real)qxy[2 * i + 1] - m.origin.y}, &ids, &w)) {
            out[i] = std::nan("");
            continue;
        }
        real z = 0;
        for (size_t k = 0; k < ids.size(); ++k) z += w[k] * (real)values[ids[k]];
        out[i] = (double)z;
    }
}

// Same values with walk-based point location (CPU baseline of bench.py's raster leg; consecutive queries should be
// close to each other).  The time spent building the mesh is kept apart: fo_nn_last_mesh_seconds().
void fo_nn_interpolate_walk(uint32_t n, const double* xy, uint32_t nt, const uint32_t* tri, const double* values,
                            uint32_t nq, const double* qxy, double* out) {
    const auto t0 = std::chrono::steady_clock::now();
    const Mesh m = build_mesh(n, xy, nt, tri);
    g_mesh_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::vector<uint32_t> ids;
    std::vector<real> w;
    uint32_t hint = 0;
    for (uint32_t i = 0; i < nq; ++i) {
        if (!query(m, Pt{(real)qxy[2 * i] - m.origin.x, (real)qxy[2 * i + 1] - m.origin.y}, &ids, &w, &hint)) {
            out[i] = std::nan("");
            continue;
        }
        real z = 0;
        for (size_t k = 0; k < ids.size(); ++k) z += w[k] * (real)values[ids[k]];
        out[i] = (double)z;
    }
}
double fo_nn_last_mesh_seconds(void) { return g_mesh_seconds; }

// Natural neighbours and weights of ONE query; returns their number (0 = None), at most `cap` are written.
uint32_t fo_nn_weights(uint32_t n, const double* xy, uint32_t nt, const uint32_t* tri, double qx, double qy,
                       uint32_t cap, uint32_t* ids_out, double* w_out) {
    const Mesh m = build_mesh(n, xy, nt, tri);
    std::vector<uint32_t> ids;
    std::vector<real> w;
    if (!query(m, Pt{(real)qx - m.origin.x, (real)qy - m.origin.y}, &ids, &w)) return 0;
    for (size_t k = 0; k < ids.size() && k < cap; ++k) { ids_out[k] = ids[k]; w_out[k] = (double)w[k]; }
    return (uint32_t)ids.size();
}

}  // extern "C"
