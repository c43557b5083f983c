"""Synthetic inputs for tests and bench in the *boundary format* of include/fastlem_b200.h.

The reference builds its model with voronoice/delaunator (src/models/surface/builder.rs:202-281); that
graph build stays on the host and is out of scope for the device path, so here it is reproduced only as
far as the hot path can observe it:
  * adjacency lists are insertion-ordered; edges are added per triangle, per half-edge, with the
    `from < to` filter of builder.rs:254-266 (so hull edges whose only half-edge has from > to are
    missing, as in the reference);
  * `dist` = Euclidean edge length (src/models/surface/sites.rs:27-34);
  * `default_outlets` = the hull cycle (builder.rs:270);
  * `areas` = a per-site cell area (inputs, not under test): barycentric dual-cell area.
Triangulation comes from scipy (Qhull) or, for very large N, from a jittered lattice split into triangles.
"""
import numpy as np


def random_sites(n, bound_min=(0.0, 0.0), bound_max=(100.0, 100.0), seed=0):
    rng = np.random.default_rng(seed)
    lo = np.asarray(bound_min, dtype=np.float64)
    hi = np.asarray(bound_max, dtype=np.float64)
    return lo + rng.random((n, 2)) * (hi - lo)


def _orient_ccw(pts, tri):
    a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
    cross = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
    tri = tri.copy()
    cw = cross < 0
    tri[cw, 1], tri[cw, 2] = tri[cw, 2].copy(), tri[cw, 1].copy()
    return tri, np.abs(cross) * 0.5


def model_from_triangles(pts, tri, hull_cycle=None, dedupe=False, clockwise=False):
    """Boundary-format model from CCW triangles, following builder.rs:249-270.  clockwise: orient the triangles the
    other way round before the half-edge pass (the rule `from < to` drops a hull edge whose only half-edge runs from the
    larger to the smaller index, so the orientation decides WHICH hull edges exist)."""
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    n = pts.shape[0]
    tri, tri_area = _orient_ccw(pts, np.asarray(tri, dtype=np.int64))
    if clockwise:
        tri = np.ascontiguousarray(tri[:, [0, 2, 1]])
    # half-edges in (triangle, k) order: (a,b), (b,c), (c,a); keep from < to
    frm = tri[:, [0, 1, 2]].reshape(-1)
    to = tri[:, [1, 2, 0]].reshape(-1)
    keep = frm < to
    ea, eb = frm[keep], to[keep]
    if dedupe:  # inputs that are not proper triangulations (jittered lattice with flipped cells): keep a simple graph
        key = ea * np.int64(n) + eb
        _, first = np.unique(key, return_index=True)
        first.sort()
        ea, eb = ea[first], eb[first]
    ne = ea.size
    d = np.sqrt((pts[ea, 0] - pts[eb, 0]) ** 2 + (pts[ea, 1] - pts[eb, 1]) ** 2)
    # add_edge(a,b,w): push (b,w) to a's list and (a,w) to b's list, in edge order
    src = np.concatenate([ea, eb])
    dst = np.concatenate([eb, ea])
    seq = np.concatenate([np.arange(ne), np.arange(ne)])
    order = np.lexsort((seq, src))
    col = dst[order].astype(np.uint32)
    dist = np.concatenate([d, d])[order]
    counts = np.bincount(src, minlength=n)
    row_ptr = np.zeros(n + 1, dtype=np.uint32)
    np.cumsum(counts, out=row_ptr[1:])
    areas = np.zeros(n, dtype=np.float64)
    for k in range(3):
        areas += np.bincount(tri[:, k], weights=tri_area / 3.0, minlength=n)
    if hull_cycle is None:
        hull_cycle = _hull_cycle(pts, tri)
    return dict(n=n, row_ptr=row_ptr, col=col, dist=np.ascontiguousarray(dist), areas=areas,
                default_outlets=np.asarray(hull_cycle, dtype=np.uint32), sites=pts, triangles=tri)


def _hull_cycle(pts, tri):
    # boundary edges = directed half-edges with no opposite half-edge
    frm = tri[:, [0, 1, 2]].reshape(-1)
    to = tri[:, [1, 2, 0]].reshape(-1)
    n = pts.shape[0]
    key = frm * n + to
    rkey = to * n + frm
    boundary = ~np.isin(key, rkey)
    verts = np.unique(np.concatenate([frm[boundary], to[boundary]]))
    c = pts[verts].mean(axis=0)
    ang = np.arctan2(pts[verts, 1] - c[1], pts[verts, 0] - c[0])
    return verts[np.argsort(ang, kind="stable")]


def halfedges_from_triangles(tri, n):
    """delaunator's `halfedges` (see fastlem_b200/triangulation.py)."""
    from fastlem_b200.triangulation import halfedges_from_triangles as f
    return f(tri, n)


def triangulation_of(model):
    """(sites, triangles[3T] uint32, halfedges[3T] uint32) of a model built by this module."""
    tri = np.ascontiguousarray(model["triangles"], dtype=np.int64)
    return (model["sites"], tri.reshape(-1).astype(np.uint32), halfedges_from_triangles(tri, model["n"]))


def delaunay_model(pts, lloyd=0, bound_min=None, bound_max=None):
    """Delaunay model of `pts`; `lloyd` steps of an approximate Lloyd relaxation first (builder.rs:155-200)."""
    from scipy.spatial import Delaunay
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    for _ in range(lloyd):
        pts = _lloyd_step(pts, Delaunay(pts), bound_min, bound_max)
    dl = Delaunay(pts)
    return model_from_triangles(pts, dl.simplices)


def _lloyd_step(pts, dl, bound_min, bound_max):
    """Move interior sites to the centroid of their Voronoi cell (fan of circumcentres)."""
    tri, _ = _orient_ccw(pts, dl.simplices.astype(np.int64))
    a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
    dd = 2.0 * (a[:, 0] * (b[:, 1] - c[:, 1]) + b[:, 0] * (c[:, 1] - a[:, 1]) + c[:, 0] * (a[:, 1] - b[:, 1]))
    a2, b2, c2 = (a ** 2).sum(1), (b ** 2).sum(1), (c ** 2).sum(1)
    ux = (a2 * (b[:, 1] - c[:, 1]) + b2 * (c[:, 1] - a[:, 1]) + c2 * (a[:, 1] - b[:, 1])) / dd
    uy = (a2 * (c[:, 0] - b[:, 0]) + b2 * (a[:, 0] - c[:, 0]) + c2 * (b[:, 0] - a[:, 0])) / dd
    cc = np.stack([ux, uy], axis=1)
    if bound_min is not None:
        cc = np.clip(cc, np.asarray(bound_min), np.asarray(bound_max))
    n = pts.shape[0]
    # undirected interior edges: pair the two half-edges via sorting on the unordered key
    frm = tri[:, [0, 1, 2]].reshape(-1)
    to = tri[:, [1, 2, 0]].reshape(-1)
    tid = np.repeat(np.arange(tri.shape[0]), 3)
    lo, hi = np.minimum(frm, to), np.maximum(frm, to)
    key = lo * n + hi
    o = np.argsort(key, kind="stable")
    ks = key[o]
    pair = ks[1:] == ks[:-1]
    i0, i1 = o[:-1][pair], o[1:][pair]
    v, w = frm[i0], to[i0]
    p, q = cc[tid[i0]], cc[tid[i1]]
    on_hull = np.zeros(n, dtype=bool)
    single = np.ones(key.size, dtype=bool)
    single[i0] = False
    single[i1] = False
    on_hull[frm[single]] = True
    on_hull[to[single]] = True
    acc_a = np.zeros(n)
    acc_c = np.zeros((n, 2))
    for s in (v, w):
        ps = pts[s]
        ar = 0.5 * np.abs((p[:, 0] - ps[:, 0]) * (q[:, 1] - ps[:, 1]) - (p[:, 1] - ps[:, 1]) * (q[:, 0] - ps[:, 0]))
        cen = (ps + p + q) / 3.0
        acc_a += np.bincount(s, weights=ar, minlength=n)
        acc_c[:, 0] += np.bincount(s, weights=ar * cen[:, 0], minlength=n)
        acc_c[:, 1] += np.bincount(s, weights=ar * cen[:, 1], minlength=n)
    out = pts.copy()
    ok = (~on_hull) & (acc_a > 0)
    out[ok] = acc_c[ok] / acc_a[ok, None]
    return out


def add_edge_sites(pts, bound_min=(0.0, 0.0), bound_max=(100.0, 100.0), edge_num_x=None, edge_num_y=None):
    """TerrainModel2DBulider::add_edge_sites (builder.rs:54-131): equally spaced sites along the bounding box, appended
    after the given sites, corner order (min,min) -> (min,max) -> (max,max) -> (max,min), edge i from corner i towards
    corner i+1 at t = j / edge_num, j = 0 .. edge_num-1."""
    pts = np.asarray(pts, dtype=np.float64)
    n = pts.shape[0]
    (x0, y0), (x1, y1) = bound_min, bound_max
    corners = [(x0, y0), (x0, y1), (x1, y1), (x1, y0)]
    out = []
    for i, c in enumerate(corners):
        nxt = corners[(i + 1) % 4]
        if i % 2 == 1:
            num = edge_num_x if edge_num_x is not None else int(np.sqrt(n / (y1 - y0) * (x1 - x0)))
        else:
            num = edge_num_y if edge_num_y is not None else int(np.sqrt(n / (x1 - x0) * (y1 - y0)))
        for j in range(num):
            t = j / num
            out.append((c[0] * (1.0 - t) + nxt[0] * t, c[1] * (1.0 - t) + nxt[1] * t))
    return np.concatenate([pts, np.array(out, dtype=np.float64).reshape(-1, 2)])


def delaunay_model_with_rim(pts, n_inner, bound_min=(0.0, 0.0), bound_max=(100.0, 100.0)):
    """Delaunay model of sites whose tail (indices >= n_inner) lies exactly ON the bounding box (add_edge_sites).  Qhull
    drops collinear hull points from the triangulation; delaunator (the crate's triangulator, exact predicates) keeps
    them: every pair of consecutive rim sites is joined through a triangle with an interior apex.  That topology is the
    Delaunay triangulation of the sites with the rim bulged outwards by a hair (strictly convex position); lengths and
    areas are then taken from the TRUE coordinates."""
    from scipy.spatial import Delaunay
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    lo, hi = np.asarray(bound_min, float), np.asarray(bound_max, float)
    mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo)
    bulged = pts.copy()
    rim = bulged[n_inner:]
    u = (rim - mid) / half  # in [-1, 1]^2, on the boundary of the square
    # push every rim site outwards along its edge's normal by eps * (1 - s^2), s = position along the edge in [-1, 1]
    eps = 1e-6 * half.max()
    on_x = np.abs(np.abs(u[:, 0]) - 1.0) < 1e-12
    on_y = np.abs(np.abs(u[:, 1]) - 1.0) < 1e-12
    rim[on_x, 0] += np.sign(u[on_x, 0]) * eps * (1.0 - u[on_x, 1] ** 2)
    rim[on_y, 1] += np.sign(u[on_y, 1]) * eps * (1.0 - u[on_y, 0] ** 2)
    # clockwise triangles: the half-edges along the rim then run in the order add_edge_sites appends the rim sites, so the
    # rule `from < to` (builder.rs:255-263) KEEPS the rim edges -- the case in which their equal lengths matter
    return model_from_triangles(pts, Delaunay(bulged).simplices, clockwise=True)


def ocean_rim_outlets(model, n_inner, band=None, seed=0):
    """Outlet mask in the manner of examples/terrain_generation_advanced.rs:178-182,343-371: the `add_edge_sites` rim
    sites (indices >= n_inner) that are "ocean" seed a flood fill through neighbouring ocean sites.  Ocean = a noise field
    below a threshold, plus (band) everything within `band` of the bounding box, so the whole rim is ocean."""
    pts, n = model["sites"], model["n"]
    sea = value_noise(pts, 4.0 / 75.0, seed=seed + 7) < -0.2
    if band is not None:
        lo, hi = pts.min(axis=0), pts.max(axis=0)
        sea |= ((pts - lo) < band).any(axis=1) | ((hi - pts) < band).any(axis=1)
    rp, col = model["row_ptr"].astype(np.int64), model["col"]
    out = np.zeros(n, dtype=bool)
    stack = [i for i in range(n_inner, n) if sea[i]]
    while stack:
        i = stack.pop()
        if out[i]:
            continue
        out[i] = True
        for j in col[rp[i]:rp[i + 1]]:
            if not out[j] and sea[j]:
                stack.append(int(j))
    return out


def lattice_model(nx, ny, bound_max=(100.0, 100.0), jitter=0.35, seed=0):
    """Jittered lattice of nx*ny sites, each cell split along a randomly chosen diagonal.

    Stand-in for a relaxed Delaunay graph where Qhull is too slow (16M sites): planar, mean degree ~6,
    near-uniform spacing.  Border sites are not jittered outward so the hull is the lattice rim.
    """
    rng = np.random.default_rng(seed)
    hx, hy = bound_max[0] / (nx - 1), bound_max[1] / (ny - 1)
    gx, gy = np.meshgrid(np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64), indexing="xy")
    jx = (rng.random((ny, nx)) - 0.5) * 2 * jitter
    jy = (rng.random((ny, nx)) - 0.5) * 2 * jitter
    jx[:, 0] = jx[:, -1] = 0
    jy[0, :] = jy[-1, :] = 0
    pts = np.stack([((gx + jx) * hx).reshape(-1), ((gy + jy) * hy).reshape(-1)], axis=1)
    idx = (np.arange(ny)[:, None] * nx + np.arange(nx)[None, :])
    v00, v10 = idx[:-1, :-1].reshape(-1), idx[:-1, 1:].reshape(-1)
    v01, v11 = idx[1:, :-1].reshape(-1), idx[1:, 1:].reshape(-1)
    flip = rng.random(v00.size) < 0.5
    t1 = np.where(flip[:, None], np.stack([v00, v10, v01], 1), np.stack([v00, v10, v11], 1))
    t2 = np.where(flip[:, None], np.stack([v10, v11, v01], 1), np.stack([v00, v11, v01], 1))
    tri = np.empty((2 * v00.size, 3), dtype=np.int64)
    tri[0::2], tri[1::2] = t1, t2
    rim = np.concatenate([idx[0, :-1], idx[:-1, -1], idx[-1, :0:-1], idx[:0:-1, 0]])
    return model_from_triangles(pts, tri, hull_cycle=rim, dedupe=True)


# ------------------------------------------------------------------------------------------------
# parameter fields
# ------------------------------------------------------------------------------------------------
def value_noise(pts, freq, seed=0, octaves=5):
    """Smooth seeded fbm value noise in [-1, 1] (stand-in for the `noise` crate's Perlin fbm)."""
    rng = np.random.default_rng(seed)
    size = 256
    out = np.zeros(pts.shape[0])
    amp, tot = 1.0, 0.0
    f = freq
    for _ in range(octaves):
        table = rng.random((size, size)) * 2 - 1
        x, y = pts[:, 0] * f, pts[:, 1] * f
        x0, y0 = np.floor(x).astype(np.int64), np.floor(y).astype(np.int64)
        fx, fy = x - x0, y - y0
        sx, sy = fx * fx * (3 - 2 * fx), fy * fy * (3 - 2 * fy)
        x0 %= size; y0 %= size
        x1, y1 = (x0 + 1) % size, (y0 + 1) % size
        v = (table[y0, x0] * (1 - sx) + table[y0, x1] * sx) * (1 - sy) + \
            (table[y1, x0] * (1 - sx) + table[y1, x1] * sx) * sy
        out += v * amp
        tot += amp
        amp *= 0.5
        f *= 2.0
    return out / tot


def uniform_params(n, erodibility=1.0, uplift=1.0):
    return dict(base=np.zeros(n), erodibility=np.full(n, float(erodibility)), uplift=np.full(n, float(uplift)),
                max_slope=None, is_outlet=np.zeros(n, dtype=bool))


def advanced_params(model, seed=0, ocean_level=-0.25):
    """terrain_generation_advanced.rs:136-210 style: noise-driven erodibility and an ocean-mask outlet set
    (sites below `ocean_level` in a second noise field, flood-filled from the hull)."""
    pts, n = model["sites"], model["n"]
    k = np.abs(value_noise(pts, 1.0 / 75.0 * 8, seed=seed)) * 4.0 + 0.1
    land = value_noise(pts, 1.0 / 75.0 * 4, seed=seed + 1000)
    sea = land < ocean_level
    # flood from hull through `sea` sites
    rp, col = model["row_ptr"].astype(np.int64), model["col"]
    is_out = np.zeros(n, dtype=bool)
    hull = model["default_outlets"].astype(np.int64)
    is_out[hull] = True
    frontier = hull
    while frontier.size:
        starts, ends = rp[frontier], rp[frontier + 1]
        cnt = ends - starts
        offs = np.repeat(starts - np.cumsum(np.concatenate([[0], cnt[:-1]])), cnt) + np.arange(cnt.sum())
        nb = np.unique(col[offs].astype(np.int64))
        nb = nb[sea[nb] & ~is_out[nb]]
        is_out[nb] = True
        frontier = nb
    p = uniform_params(n)
    p["erodibility"] = k
    p["is_outlet"] = is_out
    return p


def outlets_for(model, params):
    """generator.rs:120-132: explicit outlets ascending, else the model's default (hull) outlets."""
    idx = np.nonzero(params["is_outlet"])[0].astype(np.uint32)
    return idx if idx.size else model["default_outlets"]
