"""Host-side Delaunay triangulation in delaunator's layout, for the Python mirror of TerrainInterpolator2D.

The reference's interpolator builds its own triangulation on the host (`naturalneighbor::Interpolator::new(sites)`,
src/models/surface/interpolator.rs:11-15); the graph build stays on the host here too (north_star).  The Rust shim
uses `delaunator::triangulate` directly; this module gives the Python mirror the same arrays from
scipy's Qhull wrapper.
"""
import numpy as np

EMPTY = 0xFFFFFFFF


def halfedges_from_triangles(triangles, n_sites):
    """delaunator's `halfedges` for consistently oriented triangles (T x 3): for half-edge e = 3t+k
    (tri[t,k] -> tri[t,(k+1)%3]) the index of the opposite half-edge, 0xFFFFFFFF on the hull."""
    tri = np.asarray(triangles, dtype=np.int64).reshape(-1, 3)
    frm = tri.reshape(-1)
    to = tri[:, [1, 2, 0]].reshape(-1)
    he = np.full(frm.size, EMPTY, dtype=np.uint32)
    if frm.size == 0:
        return he
    key = frm * np.int64(n_sites) + to
    rkey = to * np.int64(n_sites) + frm
    order = np.argsort(key, kind="stable")
    skey = key[order]
    pos = np.minimum(np.searchsorted(skey, rkey), key.size - 1)
    hit = skey[pos] == rkey
    he[hit] = order[pos[hit]].astype(np.uint32)
    return he


def orient_ccw(sites, triangles):
    """Counter-clockwise copies of the triangles (T x 3)."""
    pts = np.asarray(sites, dtype=np.float64).reshape(-1, 2)
    tri = np.array(triangles, dtype=np.int64).reshape(-1, 3)
    a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
    cw = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]) < 0
    tri[cw, 1], tri[cw, 2] = tri[cw, 2].copy(), tri[cw, 1].copy()
    return tri


def delaunay(sites):
    """(triangles[3T] uint32, halfedges[3T] uint32) of the Delaunay triangulation of `sites` (n x 2)."""
    from scipy.spatial import Delaunay
    pts = np.ascontiguousarray(sites, dtype=np.float64).reshape(-1, 2)
    # Qhull lifts to x^2 + y^2 in double precision: centre the points first, or data in large absolute coordinates
    # comes back with in-circle decisions that are off by far more than rounding
    centred = pts - pts.mean(axis=0) if pts.shape[0] else pts
    tri = orient_ccw(pts, Delaunay(centred).simplices)
    return tri.reshape(-1).astype(np.uint32), halfedges_from_triangles(tri, pts.shape[0])
