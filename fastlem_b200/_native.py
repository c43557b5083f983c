"""ctypes binding of include/fastlem_b200.h.

The product library is fastlem_b200/_lib/libfastlem_b200.so (nvcc, sm_100a).  There is no CPU fallback:
if the library is missing, or no CUDA device can be opened, this module raises.  (tests/ may pass an
explicit path to the FL_EMU host build to check solver logic on the CPU tier; nothing else does.)
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libfastlem_b200.so")

UNTIL_STABLE = 0xFFFFFFFF
OK, E_INVALID, E_STATE, E_CUDA, E_NOMEM = 0, -1, -2, -3, -4

STAGE = dict(receivers=0, receivers_initial=1, labels_initial=2, labels=3, depth=4, drainage_area=5,
             response_time=6, elevation=7, flood_rank=8)
_STAGE_DTYPE = {0: np.uint32, 1: np.uint32, 2: np.uint32, 3: np.uint32, 4: np.uint32, 5: np.float64,
                6: np.float64, 7: np.float64, 8: np.uint32}

# every symbol include/fastlem_b200.h declares
SYMBOLS = ["fastlem_create", "fastlem_destroy", "fastlem_last_error", "fastlem_get_device", "fastlem_set_graph",
           "fastlem_set_parameters", "fastlem_generate", "fastlem_run", "fastlem_download", "fastlem_download_to_device",
           "fastlem_set_option", "fastlem_get_stats", "fastlem_debug_fetch", "fastlem_version", "fastlem_trim_memory",
           "fastlem_host_initial_elevations", "fastlem_host_tan_max_slope", "fastlem_host_graph_from_triangles",
           "fastlem_interp_create", "fastlem_interp_destroy", "fastlem_interp_last_error", "fastlem_interp_set_values",
           "fastlem_interp_set_values_device", "fastlem_interp_set_values_from", "fastlem_interp_points",
           "fastlem_interp_raster", "fastlem_interp_raster_device", "fastlem_interp_get_stats"]


class Stats(ctypes.Structure):
    _fields_ = [("iterations", ctypes.c_uint32), ("lake_iterations", ctypes.c_uint32),
                ("depth_first", ctypes.c_uint32), ("depth_last", ctypes.c_uint32),
                ("kernel_launches", ctypes.c_uint64), ("ms_run", ctypes.c_double), ("ms_upload", ctypes.c_double),
                ("ms_flood_rank", ctypes.c_double), ("ms_download", ctypes.c_double),
                ("ms_receivers", ctypes.c_double), ("ms_labels", ctypes.c_double), ("ms_lakes", ctypes.c_double),
                ("ms_order", ctypes.c_double), ("ms_area", ctypes.c_double), ("ms_elevation", ctypes.c_double),
                ("n_receivers", ctypes.c_uint64), ("n_labels", ctypes.c_uint64), ("n_lakes", ctypes.c_uint64),
                ("n_order", ctypes.c_uint64), ("n_area", ctypes.c_uint64), ("n_elevation", ctypes.c_uint64),
                ("rebuilds", ctypes.c_uint32), ("path_levels", ctypes.c_uint32), ("paths", ctypes.c_uint32),
                ("incremental_iterations", ctypes.c_uint32),
                ("flood_on_device", ctypes.c_uint32), ("outlet_ranks_on_device", ctypes.c_uint32),
                ("ms_kernel", ctypes.c_double * 8), ("n_kernel", ctypes.c_uint64 * 8)]

    KERNELS = ("k_receivers_bulk", "k_area_flow", "k_incr_start", "k_area_flow_long", "k_elev_plan", "k_elev_top",
               "k_elev_low", "rebuild")  # FASTLEM_K_*

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in ("ms_kernel", "n_kernel")}
        d["kernels"] = {name: {"ms": self.ms_kernel[i], "launches": int(self.n_kernel[i])}
                        for i, name in enumerate(self.KERNELS)}
        return d


class InterpStats(ctypes.Structure):
    _fields_ = [("ms_setup", ctypes.c_double), ("ms_query_kernel", ctypes.c_double), ("queries", ctypes.c_uint64),
                ("kernel_launches", ctypes.c_uint64), ("grid_x", ctypes.c_uint32), ("grid_y", ctypes.c_uint32),
                ("grid_passes", ctypes.c_uint32), ("clockwise", ctypes.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Raster(ctypes.Structure):
    """fastlem_raster: pixel (col,row) -> x = span_x*((col+pixel_offset)/width)+x0, y likewise."""
    _fields_ = [("x0", ctypes.c_double), ("y0", ctypes.c_double), ("span_x", ctypes.c_double),
                ("span_y", ctypes.c_double), ("pixel_offset", ctypes.c_double), ("width", ctypes.c_uint32),
                ("height", ctypes.c_uint32), ("row_begin", ctypes.c_uint32), ("row_end", ctypes.c_uint32)]


class FastlemError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fastlem_b200 error {code}: {msg}")
        self.code = code


_libs = {}


def load(path=None):
    path = path or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise ImportError(
            f"fastlem_b200: native library not found at {path}. Build it with `python -m fastlem_b200.build` "
            "(needs nvcc). There is no CPU fallback.")
    lib = ctypes.CDLL(path)
    vp, u32, u32p, f64p = ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32), \
        ctypes.POINTER(ctypes.c_double)
    lib.fastlem_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int]
    lib.fastlem_destroy.argtypes = [vp]
    lib.fastlem_destroy.restype = None
    lib.fastlem_last_error.argtypes = [vp]
    lib.fastlem_last_error.restype = ctypes.c_char_p
    lib.fastlem_get_device.argtypes = [vp]
    lib.fastlem_set_graph.argtypes = [vp, u32, u32p, u32p, f64p, f64p]
    lib.fastlem_set_parameters.argtypes = [vp, f64p, f64p, f64p, f64p, u32p, u32]
    lib.fastlem_generate.argtypes = [vp, u32, f64p, u32p]
    lib.fastlem_run.argtypes = [vp, u32, u32p]
    lib.fastlem_download.argtypes = [vp, f64p]
    lib.fastlem_download_to_device.argtypes = [vp, vp]
    lib.fastlem_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_int64]
    lib.fastlem_get_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    lib.fastlem_debug_fetch.argtypes = [vp, ctypes.c_int, vp, ctypes.c_size_t]
    lib.fastlem_version.restype = ctypes.c_char_p
    lib.fastlem_trim_memory.argtypes = [ctypes.c_int]
    lib.fastlem_trim_memory.restype = ctypes.c_int
    lib.fastlem_host_initial_elevations.argtypes = [u32, f64p, f64p]
    lib.fastlem_host_initial_elevations.restype = None
    lib.fastlem_host_tan_max_slope.argtypes = [u32, f64p, f64p]
    lib.fastlem_host_tan_max_slope.restype = None
    lib.fastlem_host_graph_from_triangles.argtypes = [u32, f64p, u32, u32p, u32p, u32p, f64p, ctypes.c_uint64,
                                                      ctypes.POINTER(ctypes.c_uint64)]
    lib.fastlem_interp_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int, u32, f64p, u32, u32p, u32p]
    lib.fastlem_interp_destroy.argtypes = [vp]
    lib.fastlem_interp_destroy.restype = None
    lib.fastlem_interp_last_error.argtypes = [vp]
    lib.fastlem_interp_last_error.restype = ctypes.c_char_p
    lib.fastlem_interp_set_values.argtypes = [vp, f64p]
    lib.fastlem_interp_set_values_device.argtypes = [vp, vp]
    lib.fastlem_interp_set_values_from.argtypes = [vp, vp]
    lib.fastlem_interp_points.argtypes = [vp, u32, f64p, f64p]
    lib.fastlem_interp_raster.argtypes = [vp, ctypes.POINTER(Raster), f64p]
    lib.fastlem_interp_raster_device.argtypes = [vp, ctypes.POINTER(Raster), vp]
    lib.fastlem_interp_get_stats.argtypes = [vp, ctypes.POINTER(InterpStats)]
    _libs[path] = lib
    return lib


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


class Context:
    """One fastlem_ctx (the C ABI copies everything it is handed; no array has to outlive the call)."""

    def __init__(self, device=0, lib_path=None):
        self._lib = load(lib_path)
        self._h = ctypes.c_void_p()
        rc = self._lib.fastlem_create(ctypes.byref(self._h), int(device))
        if rc != OK:
            self._h = None
            raise FastlemError(rc, f"cannot create context on CUDA device {device} (no CPU fallback)")
        self._keep = {}
        self.n = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fastlem_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != OK:
            raise FastlemError(rc, self._lib.fastlem_last_error(self._h).decode())

    def version(self):
        return self._lib.fastlem_version().decode()

    def set_option(self, name, value):
        self._ck(self._lib.fastlem_set_option(self._h, name.encode(), int(value)))

    def set_graph(self, row_ptr, col, dist, areas):
        rp, c, d, a = _u32(row_ptr), _u32(col), _f64(dist), _f64(areas)
        n = rp.size - 1
        if a.size != n or c.size != d.size:
            raise ValueError("set_graph: inconsistent array sizes")
        self._keep["graph"] = (rp, c, d, a)
        self._ck(self._lib.fastlem_set_graph(self._h, n, _p(rp, ctypes.c_uint32), _p(c, ctypes.c_uint32),
                                             _p(d, ctypes.c_double), _p(a, ctypes.c_double)))
        self.n = n

    def set_parameters(self, initial_elevation, erodibility, uplift_rate, tan_max_slope, outlets):
        e0, k, u = _f64(initial_elevation), _f64(erodibility), _f64(uplift_rate)
        t = None if tan_max_slope is None else _f64(tan_max_slope)
        o = _u32(outlets)
        for arr in (e0, k, u) + (() if t is None else (t,)):
            if arr.size != self.n:
                raise ValueError("set_parameters: parameter arrays must have one entry per site")
        self._keep["params"] = (e0, k, u, t, o)
        self._ck(self._lib.fastlem_set_parameters(self._h, _p(e0, ctypes.c_double), _p(k, ctypes.c_double),
                                                  _p(u, ctypes.c_double), _p(t, ctypes.c_double),
                                                  _p(o, ctypes.c_uint32), o.size))

    def generate(self, max_iteration=None, out=None):
        out = np.empty(self.n, dtype=np.float64) if out is None else out
        it = ctypes.c_uint32(0)
        mi = UNTIL_STABLE if max_iteration is None else int(max_iteration)
        self._ck(self._lib.fastlem_generate(self._h, mi, _p(out, ctypes.c_double), ctypes.byref(it)))
        return out, it.value

    def run(self, max_iteration=None):
        it = ctypes.c_uint32(0)
        mi = UNTIL_STABLE if max_iteration is None else int(max_iteration)
        self._ck(self._lib.fastlem_run(self._h, mi, ctypes.byref(it)))
        return it.value

    def download(self, out=None):
        out = np.empty(self.n, dtype=np.float64) if out is None else out
        self._ck(self._lib.fastlem_download(self._h, _p(out, ctypes.c_double)))
        return out

    def download_to_device(self, device_ptr):
        """Copy the elevations into a device buffer (e.g. torch_tensor.data_ptr()) of n float64 on this device."""
        self._ck(self._lib.fastlem_download_to_device(self._h, ctypes.c_void_p(int(device_ptr))))

    def stats(self):
        s = Stats()
        self._ck(self._lib.fastlem_get_stats(self._h, ctypes.byref(s)))
        return s.as_dict()

    def fetch(self, stage):
        code = STAGE[stage] if isinstance(stage, str) else int(stage)
        out = np.empty(self.n, dtype=_STAGE_DTYPE[code])
        self._ck(self._lib.fastlem_debug_fetch(self._h, code, out.ctypes.data_as(ctypes.c_void_p), out.nbytes))
        return out


class Interpolator:
    """One fastlem_interp: natural-neighbour interpolation over a Delaunay triangulation in delaunator's layout
    (triangles[3T], halfedges[3T] with 0xFFFFFFFF on the hull)."""

    def __init__(self, sites, triangles, halfedges, device=0, lib_path=None):
        self._lib = load(lib_path)
        self._h = ctypes.c_void_p()
        xy = _f64(sites).reshape(-1)
        tri, he = _u32(triangles).reshape(-1), _u32(halfedges).reshape(-1)
        if xy.size % 2 or tri.size % 3 or tri.size != he.size:
            raise ValueError("Interpolator: inconsistent array sizes")
        self.n = xy.size // 2
        rc = self._lib.fastlem_interp_create(ctypes.byref(self._h), int(device), self.n, _p(xy, ctypes.c_double),
                                             tri.size // 3, _p(tri, ctypes.c_uint32), _p(he, ctypes.c_uint32))
        if rc != OK:
            self._h = None
            raise FastlemError(rc, "cannot create interpolator on CUDA device %d (details on stderr; no CPU fallback)"
                               % device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fastlem_interp_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != OK:
            raise FastlemError(rc, self._lib.fastlem_interp_last_error(self._h).decode())

    def set_values(self, values):
        v = _f64(values)
        if v.size != self.n:
            raise ValueError("set_values: one value per site")
        self._ck(self._lib.fastlem_interp_set_values(self._h, _p(v, ctypes.c_double)))

    def set_values_device(self, device_ptr):
        self._ck(self._lib.fastlem_interp_set_values_device(self._h, ctypes.c_void_p(int(device_ptr))))

    def set_values_from(self, ctx):
        """Take the elevations of a solver Context that has run, device to device."""
        self._ck(self._lib.fastlem_interp_set_values_from(self._h, ctx._h))

    def points(self, points_xy, out=None):
        q = _f64(points_xy).reshape(-1)
        if q.size % 2:
            raise ValueError("points: x y pairs expected")
        nq = q.size // 2
        out = np.empty(nq, dtype=np.float64) if out is None else out
        self._ck(self._lib.fastlem_interp_points(self._h, nq, _p(q, ctypes.c_double), _p(out, ctypes.c_double)))
        return out

    @staticmethod
    def raster_desc(width, height, x0, y0, span_x, span_y, pixel_offset=0.0, row_begin=0, row_end=None):
        return Raster(float(x0), float(y0), float(span_x), float(span_y), float(pixel_offset), int(width), int(height),
                      int(row_begin), int(height if row_end is None else row_end))

    def raster(self, desc, out=None):
        rows = desc.row_end - desc.row_begin
        out = np.empty((max(rows, 0), desc.width), dtype=np.float64) if out is None else out
        self._ck(self._lib.fastlem_interp_raster(self._h, ctypes.byref(desc), _p(out, ctypes.c_double)))
        return out

    def raster_device(self, desc, device_ptr):
        self._ck(self._lib.fastlem_interp_raster_device(self._h, ctypes.byref(desc), ctypes.c_void_p(int(device_ptr))))

    def stats(self):
        s = InterpStats()
        self._ck(self._lib.fastlem_interp_get_stats(self._h, ctypes.byref(s)))
        return s.as_dict()


def host_initial_elevations(base_elevation, lib_path=None):
    """generator.rs:134-138 on the host: base + StdRng::seed_from_u64(0).gen::<f64>() * f64::EPSILON."""
    base = _f64(base_elevation)
    out = np.empty_like(base)
    load(lib_path).fastlem_host_initial_elevations(base.size, _p(base, ctypes.c_double), _p(out, ctypes.c_double))
    return out


def host_tan_max_slope(max_slope, lib_path=None):
    """generator.rs:194 on the host: tan(max_slope) with libm (what Rust's f64::tan calls); NaN (None) stays NaN."""
    ms = _f64(max_slope)
    out = np.empty_like(ms)
    load(lib_path).fastlem_host_tan_max_slope(ms.size, _p(ms, ctypes.c_double), _p(out, ctypes.c_double))
    return out


def host_graph_from_triangles(sites, triangles, lib_path=None):
    """builder.rs:252-268 on the host: (row_ptr, col, dist) in the boundary format of set_graph, rows in
    neighbors_of order, from the builder's triangle array (3T site indices)."""
    lib = load(lib_path)
    xy, tri = _f64(sites).reshape(-1), _u32(triangles).reshape(-1)
    n, nt = xy.size // 2, tri.size // 3
    row_ptr = np.empty(n + 1, dtype=np.uint32)
    nnz = ctypes.c_uint64(0)
    rc = lib.fastlem_host_graph_from_triangles(n, _p(xy, ctypes.c_double), nt, _p(tri, ctypes.c_uint32),
                                               _p(row_ptr, ctypes.c_uint32), None, None, 0, ctypes.byref(nnz))
    if rc != OK:
        raise FastlemError(rc, "graph_from_triangles: invalid triangulation")
    col = np.empty(nnz.value, dtype=np.uint32)
    dist = np.empty(nnz.value, dtype=np.float64)
    rc = lib.fastlem_host_graph_from_triangles(n, _p(xy, ctypes.c_double), nt, _p(tri, ctypes.c_uint32),
                                               _p(row_ptr, ctypes.c_uint32), _p(col, ctypes.c_uint32),
                                               _p(dist, ctypes.c_double), nnz.value, ctypes.byref(nnz))
    if rc != OK:
        raise FastlemError(rc, "graph_from_triangles failed")
    return row_ptr, col, dist
