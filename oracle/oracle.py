"""ctypes wrapper around the CPU oracle (TEST INFRASTRUCTURE ONLY).

Loads oracle/libfastlem_oracle.so (built by oracle/Makefile from fastlem_oracle.cpp, the
single-threaded restatement of /root/reference src/lem/{generator,stream_tree,drainage_basin}.rs, and from
nn_oracle.cpp, the natural-neighbour interpolation behind Terrain2D::get_elevation).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfastlem_oracle.so")
NONE = 0xFFFFFFFF

_u32p = ctypes.POINTER(ctypes.c_uint32)
_f64p = ctypes.POINTER(ctypes.c_double)
_intp = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("fastlem_oracle.cpp", "nn_oracle.cpp", "Makefile")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.fo_generate.restype = ctypes.c_uint32
        _lib.fo_iterate_once.restype = ctypes.c_int
        _lib.fo_nn_weights.restype = ctypes.c_uint32
        _lib.fo_nn_last_mesh_seconds.restype = ctypes.c_double
    return _lib


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def chacha_block(key_words, counter, double_rounds):
    key = _u32(key_words)
    out = np.zeros(16, dtype=np.uint32)
    lib().fo_chacha_block(_p(key, _u32p), ctypes.c_uint64(counter), ctypes.c_int(double_rounds), _p(out, _u32p))
    return out


def seed_from_u64(state):
    out = np.zeros(32, dtype=np.uint8)
    lib().fo_seed_from_u64(ctypes.c_uint64(state), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    return out


def gen_f64(seed, n):
    out = np.zeros(n, dtype=np.float64)
    lib().fo_gen_f64(ctypes.c_uint64(seed), ctypes.c_uint32(n), _p(out, _f64p))
    return out


def initial_elevations(base_elevation):
    base = _f64(base_elevation)
    out = np.empty_like(base)
    lib().fo_initial_elevations(ctypes.c_uint32(base.size), _p(base, _f64p), _p(out, _f64p))
    return out


def random_sites(n, bound_min, bound_max, seed_byte=0):
    out = np.empty((n, 2), dtype=np.float64)
    lib().fo_random_sites(ctypes.c_uint32(n), ctypes.c_uint8(seed_byte), ctypes.c_double(bound_min[0]),
                          ctypes.c_double(bound_min[1]), ctypes.c_double(bound_max[0]),
                          ctypes.c_double(bound_max[1]), _p(out, _f64p))
    return out


def _graph_args(m):
    rp, col, dist = _u32(m["row_ptr"]), _u32(m["col"]), _f64(m["dist"])
    n = rp.size - 1
    return n, rp, col, dist


def flood_order(m, outlets):
    n, rp, col, dist = _graph_args(m)
    outlets = _u32(outlets)
    out = np.empty(n, dtype=np.uint32)
    lib().fo_flood_order(ctypes.c_uint32(n), _p(rp, _u32p), _p(col, _u32p), _p(dist, _f64p), _p(outlets, _u32p),
                         ctypes.c_uint32(outlets.size), _p(out, _u32p))
    return out


def stream_tree(m, elevations, outlets):
    n, rp, col, dist = _graph_args(m)
    outlets = _u32(outlets)
    e = _f64(elevations)
    nxt = np.empty(n, dtype=np.uint32)
    nxt0 = np.empty(n, dtype=np.uint32)
    sub = np.empty(n, dtype=np.uint32)
    hl = ctypes.c_int(0)
    lib().fo_stream_tree(ctypes.c_uint32(n), _p(rp, _u32p), _p(col, _u32p), _p(dist, _f64p), _p(e, _f64p),
                         _p(outlets, _u32p), ctypes.c_uint32(outlets.size), _p(nxt, _u32p), _p(nxt0, _u32p),
                         _p(sub, _u32p), ctypes.byref(hl))
    return dict(next=nxt, next_initial=nxt0, subroot=sub, has_lake=bool(hl.value))


def iterate_once(m, erodibility, uplift_rate, max_slope, outlets, elevations):
    """One loop body on a copy of `elevations`; returns dict with every stage."""
    n, rp, col, dist = _graph_args(m)
    areas = _f64(m["areas"])
    outlets = _u32(outlets)
    e = _f64(elevations).copy()
    k, u = _f64(erodibility), _f64(uplift_rate)
    ms = None if max_slope is None else _f64(max_slope)
    nxt = np.empty(n, dtype=np.uint32)
    nxt0 = np.empty(n, dtype=np.uint32)
    sub = np.empty(n, dtype=np.uint32)
    order = np.empty(n, dtype=np.uint32)
    A = np.empty(n, dtype=np.float64)
    rt = np.empty(n, dtype=np.float64)
    hl = ctypes.c_int(0)
    changed = lib().fo_iterate_once(ctypes.c_uint32(n), _p(rp, _u32p), _p(col, _u32p), _p(dist, _f64p),
                                    _p(areas, _f64p), _p(k, _f64p), _p(u, _f64p), _p(ms, _f64p), _p(outlets, _u32p),
                                    ctypes.c_uint32(outlets.size), _p(e, _f64p), _p(nxt, _u32p), _p(nxt0, _u32p),
                                    _p(sub, _u32p), ctypes.byref(hl), _p(A, _f64p), _p(rt, _f64p), _p(order, _u32p))
    return dict(elevations=e, next=nxt, next_initial=nxt0, subroot=sub, has_lake=bool(hl.value), drainage=A,
                response=rt, order=order, changed=bool(changed))


def generate(m, erodibility, uplift_rate, max_slope, outlets, initial, max_iteration=None):
    """generator.rs:140-210 from `initial` (= base + noise). Returns (elevations, iterations)."""
    n, rp, col, dist = _graph_args(m)
    areas = _f64(m["areas"])
    outlets = _u32(outlets)
    e = _f64(initial).copy()
    k, u = _f64(erodibility), _f64(uplift_rate)
    ms = None if max_slope is None else _f64(max_slope)
    mi = 0xFFFFFFFF if max_iteration is None else int(max_iteration)
    it = lib().fo_generate(ctypes.c_uint32(n), _p(rp, _u32p), _p(col, _u32p), _p(dist, _f64p), _p(areas, _f64p),
                           _p(k, _f64p), _p(u, _f64p), _p(ms, _f64p), _p(outlets, _u32p),
                           ctypes.c_uint32(outlets.size), ctypes.c_uint32(mi), _p(e, _f64p))
    return e, int(it)


# ------------------------------------------------------------------------------------------------
# Terrain2D::get_elevation (nn_oracle.cpp)
# ------------------------------------------------------------------------------------------------
def nn_last_mesh_seconds():
    return float(lib().fo_nn_last_mesh_seconds())


def nn_interpolate(sites, triangles, values, queries, walk=False):
    """terrain.rs:36-38 / interpolator.rs:17-27 for a batch of points; NaN = None.
    walk=True: point location by walking from the previous query's triangle instead of brute force (same values)."""
    xy, tri, v, q = _f64(sites).reshape(-1), _u32(triangles).reshape(-1), _f64(values), _f64(queries).reshape(-1)
    out = np.empty(q.size // 2, dtype=np.float64)
    f = lib().fo_nn_interpolate_walk if walk else lib().fo_nn_interpolate
    f(ctypes.c_uint32(xy.size // 2), _p(xy, _f64p), ctypes.c_uint32(tri.size // 3), _p(tri, _u32p),
                            _p(v, _f64p), ctypes.c_uint32(out.size), _p(q, _f64p), _p(out, _f64p))
    return out


def nn_weights(sites, triangles, x, y, cap=256):
    """Natural neighbours and Sibson weights of one query: (ids, weights), empty for None."""
    xy, tri = _f64(sites).reshape(-1), _u32(triangles).reshape(-1)
    ids = np.empty(cap, dtype=np.uint32)
    w = np.empty(cap, dtype=np.float64)
    k = lib().fo_nn_weights(ctypes.c_uint32(xy.size // 2), _p(xy, _f64p), ctypes.c_uint32(tri.size // 3), _p(tri, _u32p),
                            ctypes.c_double(x), ctypes.c_double(y), ctypes.c_uint32(cap), _p(ids, _u32p), _p(w, _f64p))
    return ids[:k].copy(), w[:k].copy()
