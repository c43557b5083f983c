"""fastlem_b200 -- B200-native `TerrainGenerator::generate()` for TadaTeruki/fastlem.

Python host-side mirror of the reference's public interface for the one path this package replaces
(the C++ mirror with the same names is fastlem_b200/host/fastlem.hpp; the Rust shim is in INTEGRATION.md):

    reference (Rust)                                     here
    ---------------------------------------------------  ------------------------------------------
    core::parameters::TopographicalParameters            TopographicalParameters   (parameters.rs:24-69)
    models::surface::sites::Site2D                       Site2D                    (sites.rs)
    models::surface::model::TerrainModel2D (trait Model) TerrainModel2D            (model.rs:41-69, traits.rs:13-20)
    models::surface::terrain::Terrain2D                  Terrain2D                 (terrain.rs:8-39)
    models::surface::interpolator::TerrainInterpolator2D TerrainInterpolator2D     (interpolator.rs:6-28)
    lem::generator::TerrainGenerator                     TerrainGenerator          (generator.rs:36-213)
    lem::generator::GenerationError                      GenerationError + variants (generator.rs:18-26)

`generate()` packs the model into the C ABI's boundary format and runs the whole loop on the GPU through
include/fastlem_b200.h.  There is no CPU implementation in this package.
"""
import math

import numpy as np

from . import _native

__all__ = ["Site2D", "TopographicalParameters", "ParameterArrays", "TerrainModel2D", "Terrain2D",
           "TerrainInterpolator2D", "TerrainGenerator", "GenerationError", "InvalidNumberOfParameters", "ParametersNotSet", "ModelNotSet"]


class GenerationError(Exception):
    """generator.rs:18-26"""


class InvalidNumberOfParameters(GenerationError):
    def __init__(self):
        super().__init__("The number of topographical parameters must be equal to the number of sites")


class ParametersNotSet(GenerationError):
    def __init__(self):
        super().__init__("You must set topographical parameters before generating terrain")


class ModelNotSet(GenerationError):
    def __init__(self):
        super().__init__("You must set `TerrainModel` before generating terrain")


class Site2D:
    __slots__ = ("x", "y")

    def __init__(self, x=0.0, y=0.0):
        self.x, self.y = float(x), float(y)

    def distance(self, other):
        return math.sqrt(self.squared_distance(other))

    def squared_distance(self, other):
        return (self.x - other.x) ** 2 + (self.y - other.y) ** 2


class TopographicalParameters:
    """parameters.rs:24-69: defaults 0 / 1 / 1 / False / None, chained setters."""
    __slots__ = ("base_elevation", "erodibility", "uplift_rate", "is_outlet", "max_slope")

    def __init__(self):
        self.base_elevation = 0.0
        self.erodibility = 1.0
        self.uplift_rate = 1.0
        self.is_outlet = False
        self.max_slope = None

    @classmethod
    def default(cls):
        return cls()

    def set_base_elevation(self, v):
        self.base_elevation = float(v)
        return self

    def set_erodibility(self, v):
        self.erodibility = float(v)
        return self

    def set_uplift_rate(self, v):
        self.uplift_rate = float(v)
        return self

    def set_is_outlet(self, v):
        self.is_outlet = bool(v)
        return self

    def set_max_slope(self, v):
        self.max_slope = None if v is None else float(v)
        return self


class ParameterArrays:
    """Structure-of-arrays form of Vec<TopographicalParameters> for large models (what the C ABI takes).
    max_slope: radians, NaN = None (or None for "None everywhere")."""

    def __init__(self, base_elevation, erodibility, uplift_rate, is_outlet=None, max_slope=None):
        self.base_elevation = np.ascontiguousarray(base_elevation, dtype=np.float64)
        self.erodibility = np.ascontiguousarray(erodibility, dtype=np.float64)
        self.uplift_rate = np.ascontiguousarray(uplift_rate, dtype=np.float64)
        n = self.base_elevation.size
        self.is_outlet = np.zeros(n, dtype=bool) if is_outlet is None else np.ascontiguousarray(is_outlet, dtype=bool)
        self.max_slope = None if max_slope is None else np.ascontiguousarray(max_slope, dtype=np.float64)

    def __len__(self):
        return self.base_elevation.size

    @classmethod
    def from_list(cls, params):
        n = len(params)
        ms = None
        if any(p.max_slope is not None for p in params):
            ms = np.array([np.nan if p.max_slope is None else p.max_slope for p in params], dtype=np.float64)
        return cls(np.fromiter((p.base_elevation for p in params), np.float64, n),
                   np.fromiter((p.erodibility for p in params), np.float64, n),
                   np.fromiter((p.uplift_rate for p in params), np.float64, n),
                   np.fromiter((p.is_outlet for p in params), bool, n), ms)


class TerrainModel2D:
    """model.rs:18-69.  The graph is held as the CSR the C ABI takes: row i = graph.neighbors_of(i) in order."""

    def __init__(self, sites, areas, row_ptr, col, dist, default_outlets, triangles=None, halfedges=None):
        # triangles / halfedges: the builder's Delaunay triangulation in delaunator's layout, if it is at hand; the
        # interpolator of the resulting Terrain2D reuses it instead of triangulating a second time
        self._triangulation = None
        if triangles is not None:
            from . import triangulation as _tr
            tri = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1)
            he = _tr.halfedges_from_triangles(tri, len(areas)) if halfedges is None else \
                np.ascontiguousarray(halfedges, dtype=np.uint32).reshape(-1)
            self._triangulation = (tri, he)
        self._sites = np.ascontiguousarray(sites, dtype=np.float64).reshape(-1, 2)
        self._areas = np.ascontiguousarray(areas, dtype=np.float64)
        self._row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
        self._col = np.ascontiguousarray(col, dtype=np.uint32)
        self._dist = np.ascontiguousarray(dist, dtype=np.float64)
        self._default_outlets = np.ascontiguousarray(default_outlets, dtype=np.uint32)

    @classmethod
    def from_workload(cls, m):
        return cls(m["sites"], m["areas"], m["row_ptr"], m["col"], m["dist"], m["default_outlets"],
                   triangles=m.get("triangles"))

    def num(self):  # model.rs:42-44: graph.order()
        return self._row_ptr.size - 1

    def sites(self):
        return self._sites

    def areas(self):
        return self._areas

    def default_outlets(self):
        return self._default_outlets

    def graph(self):
        return self._row_ptr, self._col, self._dist

    def create_terrain_from_result(self, elevations, device=0, lib_path=None):  # model.rs:62-68
        sites = self._sites.copy()
        own = np.array(elevations, dtype=np.float64, copy=True)
        own.setflags(write=False)  # (the reference's Terrain2D owns its elevations; the interpolator caches the upload)
        return Terrain2D(sites, own, TerrainInterpolator2D(sites, self._triangulation, device, lib_path))


class TerrainInterpolator2D:
    """interpolator.rs:6-28.  `new(sites)` is lazy: the reference triangulates the sites a second time inside
    generate() (model.rs:62-68); here the device interpolator is created on the first query, from the builder's
    triangulation when the model carries one, else from a host Delaunay of the sites (triangulation.py)."""

    def __init__(self, sites, triangulation=None, device=0, lib_path=None):
        self._sites = np.ascontiguousarray(sites, dtype=np.float64).reshape(-1, 2)
        self._triangulation = triangulation
        self._device, self._lib_path = device, lib_path
        self._native = None
        self._values_of = None
        self._has_nan = False

    @classmethod
    def new(cls, sites):
        return cls(sites)

    def _handle(self, elevations):
        if self._native is None:
            if self._triangulation is None:
                from . import triangulation as _tr
                self._triangulation = _tr.delaunay(self._sites)
            tri, he = self._triangulation
            self._native = _native.Interpolator(self._sites, tri, he, self._device, self._lib_path)
        # The reference reads the slice on every call.  The uploaded copy is reused only for an array that cannot have
        # changed in between: the same object, not writeable (Terrain2D's own elevations are made read-only).
        e = np.asarray(elevations)
        if self._values_of is not elevations or e.flags.writeable:
            self._native.set_values(elevations)
            self._values_of = elevations
            self._has_nan = bool(np.isnan(e).any())
        return self._native

    def _inside_hull(self, points_xy):
        """True where the reference returns Some(..): the same query over a constant field (NaN only outside the hull)."""
        h = self._native
        h.set_values(np.ones(self._sites.shape[0]))
        inside = ~np.isnan(h.points(points_xy))
        self._values_of = None  # the next query uploads its elevations again
        return inside

    def interpolate(self, elevations, site):  # interpolator.rs:17-27
        xy = np.array([[site.x, site.y]], dtype=np.float64)
        z = self._handle(elevations).points(xy)[0]
        if z != z:  # None outside the hull -- or Some(NaN) when the elevations themselves hold a NaN
            return float(z) if self._has_nan and self._inside_hull(xy)[0] else None
        return float(z)

    def interpolate_many(self, elevations, points_xy):
        """One call for many sites: (k, 2) array -> k values, NaN where the reference returns None."""
        return self._handle(elevations).points(points_xy)

    def raster(self, elevations, width, height, x0, y0, span_x, span_y, pixel_offset=0.0, row_begin=0, row_end=None,
               device_ptr=None):
        h = self._handle(elevations)
        desc = h.raster_desc(width, height, x0, y0, span_x, span_y, pixel_offset, row_begin, row_end)
        if device_ptr is not None:
            h.raster_device(desc, device_ptr)
            return None
        return h.raster(desc)

    def stats(self):
        return None if self._native is None else self._native.stats()

    def close(self):
        if self._native is not None:
            self._native.close()
            self._native = None
            self._values_of = None


class Terrain2D:
    """terrain.rs:8-39."""

    def __init__(self, sites, elevations, interpolator=None):
        self._sites, self._elevations = sites, elevations
        self._interpolator = interpolator if interpolator is not None else TerrainInterpolator2D(sites)

    def sites(self):
        return self._sites

    def elevations(self):
        return self._elevations

    def get_elevation(self, site):
        """terrain.rs:36-38: interpolated elevation at `site`, None outside the convex hull of the sites."""
        return self._interpolator.interpolate(self._elevations, site)

    def get_elevations(self, points_xy):
        """get_elevation for an array of points (k, 2) in one device call; NaN = None."""
        return self._interpolator.interpolate_many(self._elevations, points_xy)

    def raster(self, width, height, x0, y0, span_x, span_y, pixel_offset=0.0, row_begin=0, row_end=None,
               device_ptr=None):
        """The per-pixel get_elevation loop of the examples as one device call (include/fastlem_b200.h,
        fastlem_interp_raster): rows [row_begin, row_end) of a width x height image, NaN = None."""
        return self._interpolator.raster(self._elevations, width, height, x0, y0, span_x, span_y, pixel_offset,
                                         row_begin, row_end, device_ptr)


class TerrainGenerator:
    """generator.rs:36-213: builder + generate()."""

    def __init__(self):
        self._model = None
        self._parameters = None
        self._max_iteration = None
        self.device = 0
        self._lib_path = None  # tests only
        self.last_stats = None
        self.last_iterations = None

    @classmethod
    def default(cls):
        return cls()

    def set_model(self, model):
        self._model = model
        return self

    def set_parameters(self, parameters):
        self._parameters = parameters
        return self

    def set_max_iteration(self, max_iteration):
        self._max_iteration = int(max_iteration)
        return self

    def set_device(self, ordinal):
        self.device = int(ordinal)
        return self

    def generate(self):
        model = self._model
        if model is None:  # generator.rs:91-97
            raise ModelNotSet()
        num = model.num()
        if self._parameters is None:  # generator.rs:107-116
            raise ParametersNotSet()
        if len(self._parameters) != num:
            raise InvalidNumberOfParameters()
        p = self._parameters if isinstance(self._parameters, ParameterArrays) else \
            ParameterArrays.from_list(self._parameters)
        # generator.rs:120-132
        outlets = np.nonzero(p.is_outlet)[0].astype(np.uint32)
        if outlets.size == 0:
            outlets = model.default_outlets()
        # generator.rs:134-138
        initial = _native.host_initial_elevations(p.base_elevation, self._lib_path)
        # generator.rs:194 `max_slope.tan()`: libm's tan (np.tan's SIMD path differs by an ulp for ~0.5 % of the inputs)
        tan = None if p.max_slope is None else _native.host_tan_max_slope(p.max_slope, self._lib_path)
        row_ptr, col, dist = model.graph()
        with _native.Context(self.device, self._lib_path) as ctx:
            ctx.set_graph(row_ptr, col, dist, model.areas())
            ctx.set_parameters(initial, p.erodibility, p.uplift_rate, tan, outlets)
            elevations, it = ctx.generate(self._max_iteration)
            self.last_stats = ctx.stats()
            self.last_iterations = it
        return model.create_terrain_from_result(elevations, self.device, self._lib_path)  # generator.rs:212
