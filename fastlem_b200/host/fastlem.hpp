// fastlem.hpp -- C++ host-side mirror of the reference's public interface for the one path this library
// replaces, written above the C ABI (include/fastlem_b200.h).  Header-only; link libfastlem_b200.so.
//
//   reference (Rust, /root/reference)                       here (namespace fastlem)
//   ------------------------------------------------------  ---------------------------------------------
//   core::units (src/core/units.rs:1-23)                     Length, Elevation, ... = double; Step = uint32_t
//   core::parameters::TopographicalParameters (:24-69)       TopographicalParameters (same defaults, chained setters)
//   core::traits::Model (src/core/traits.rs:13-20)           any type with num/sites/areas/default_outlets/graph/
//                                                            create_terrain_from_result (duck-typed template)
//   models::surface::sites::Site2D (sites.rs)                Site2D
//   models::surface::model::TerrainModel2D (model.rs:18-69)  TerrainModel2D (built from a finished graph; the
//                                                            Delaunay/Lloyd builder stays outside this path)
//   models::surface::terrain::Terrain2D (terrain.rs:8-39)    Terrain2D (get_elevation on the device; + raster())
//   models::surface::interpolator::TerrainInterpolator2D     TerrainInterpolator2D (interpolator.rs:6-28; built lazily
//                                                            from the builder's Delaunay triangulation)
//   lem::generator::GenerationError (generator.rs:18-26)     GenerationError
//   lem::generator::TerrainGenerator (generator.rs:36-213)   TerrainGenerator<M, T>
//
// generate() returns Result<T, GenerationError> in Rust; here it returns fastlem::Result<T> with the same three
// validation variants plus DeviceError (the reference has no variant for that).
#pragma once

#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/fastlem_b200.h"

namespace fastlem {

using Length = double;
using Elevation = double;
using UpliftRate = double;
using Erodibility = double;
using Area = double;
using Slope = double;
using Step = std::uint32_t;

struct Site2D {
    double x = 0.0, y = 0.0;
    Length squared_distance(const Site2D& o) const { return (x - o.x) * (x - o.x) + (y - o.y) * (y - o.y); }
    Length distance(const Site2D& o) const { return std::sqrt(squared_distance(o)); }
};

// src/core/parameters.rs:24-69
class TopographicalParameters {
  public:
    Elevation base_elevation = 0.0;
    Erodibility erodibility = 1.0;
    UpliftRate uplift_rate = 1.0;
    bool is_outlet = false;
    std::optional<Slope> max_slope;

    static TopographicalParameters default_() { return TopographicalParameters(); }
    TopographicalParameters set_base_elevation(Elevation v) const { auto p = *this; p.base_elevation = v; return p; }
    TopographicalParameters set_erodibility(Erodibility v) const { auto p = *this; p.erodibility = v; return p; }
    TopographicalParameters set_uplift_rate(UpliftRate v) const { auto p = *this; p.uplift_rate = v; return p; }
    TopographicalParameters set_is_outlet(bool v) const { auto p = *this; p.is_outlet = v; return p; }
    TopographicalParameters set_max_slope(std::optional<Slope> v) const { auto p = *this; p.max_slope = v; return p; }
};

// EdgeAttributedUndirectedGraph<Length> as the hot path sees it: row i = neighbors_of(i) in order.
struct Graph {
    std::vector<std::uint32_t> row_ptr;  // n + 1
    std::vector<std::uint32_t> col;
    std::vector<Length> dist;
    std::size_t order() const { return row_ptr.empty() ? 0 : row_ptr.size() - 1; }
};

// The builder's Delaunay triangulation in delaunator's layout (what `delaunator::triangulate` returns in the crate):
// half-edge e = 3t+k runs triangles[e] -> triangles[3t+(k+1)%3]; halfedges[e] = opposite half-edge or 0xFFFFFFFF.
struct Triangulation {
    std::vector<std::uint32_t> triangles;
    std::vector<std::uint32_t> halfedges;
    bool empty() const { return triangles.empty(); }
};

// src/models/surface/interpolator.rs:6-28.  The reference triangulates the sites again inside `new`; here the
// builder's triangulation is reused and the device interpolator is created on the first query.
class TerrainInterpolator2D {
  public:
    TerrainInterpolator2D() = default;
    TerrainInterpolator2D(const std::vector<Site2D>& sites, Triangulation tri, int device = 0)
        : state_(std::make_shared<State>()) {
        state_->xy.reserve(2 * sites.size());
        for (const Site2D& s : sites) { state_->xy.push_back(s.x); state_->xy.push_back(s.y); }
        state_->tri = std::move(tri);
        state_->device = device;
    }

    // interpolator.rs:17-27: None outside the convex hull of the sites
    std::optional<Elevation> interpolate(const std::vector<Elevation>& elevations, const Site2D& site) const {
        const double q[2] = {site.x, site.y};
        double z = 0.0;
        check(fastlem_interp_points(handle(elevations), 1, q, &z));
        if (z != z) return std::nullopt;
        return z;
    }

    // rows [row_begin, row_end) of the examples' per-pixel loop in one call; NaN = None
    std::vector<Elevation> raster(const std::vector<Elevation>& elevations, const fastlem_raster& r) const {
        std::vector<Elevation> out((std::size_t)(r.row_end - r.row_begin) * r.width);
        check(fastlem_interp_raster(handle(elevations), &r, out.data()));
        return out;
    }

  private:
    struct State {
        std::vector<double> xy;
        Triangulation tri;
        int device = 0;
        fastlem_interp* h = nullptr;
        const double* values_of = nullptr;
        ~State() { fastlem_interp_destroy(h); }
    };
    fastlem_interp* handle(const std::vector<Elevation>& elevations) const {
        if (!state_ || state_->tri.empty())
            throw std::runtime_error("TerrainInterpolator2D: the model carries no triangulation");
        State& s = *state_;
        if (!s.h) {
            if (fastlem_interp_create(&s.h, s.device, (std::uint32_t)(s.xy.size() / 2), s.xy.data(),
                                      (std::uint32_t)(s.tri.triangles.size() / 3), s.tri.triangles.data(),
                                      s.tri.halfedges.data()) != FASTLEM_OK)
                throw std::runtime_error("fastlem_interp_create failed (invalid triangulation or no CUDA device; "
                                         "there is no CPU fallback)");
        }
        if (s.values_of != elevations.data()) {
            if (elevations.size() != s.xy.size() / 2) throw std::runtime_error("interpolate: one elevation per site");
            check(fastlem_interp_set_values(s.h, elevations.data()));
            s.values_of = elevations.data();
        }
        return s.h;
    }
    void check(int rc) const {
        if (rc != FASTLEM_OK) throw std::runtime_error(fastlem_interp_last_error(state_ ? state_->h : nullptr));
    }
    std::shared_ptr<State> state_;  // Terrain2D is Clone in the reference: copies share the device interpolator
};

// src/models/surface/terrain.rs:8-39
class Terrain2D {
  public:
    Terrain2D(std::vector<Site2D> sites, std::vector<Elevation> elevations,
              TerrainInterpolator2D interpolator = TerrainInterpolator2D())
        : sites_(std::move(sites)), elevations_(std::move(elevations)), interpolator_(std::move(interpolator)) {}
    const std::vector<Site2D>& sites() const { return sites_; }
    const std::vector<Elevation>& elevations() const { return elevations_; }
    // terrain.rs:36-38
    std::optional<Elevation> get_elevation(const Site2D& site) const {
        return interpolator_.interpolate(elevations_, site);
    }
    std::vector<Elevation> raster(const fastlem_raster& r) const { return interpolator_.raster(elevations_, r); }

  private:
    std::vector<Site2D> sites_;
    std::vector<Elevation> elevations_;
    TerrainInterpolator2D interpolator_;
};

// src/models/surface/model.rs:18-69
class TerrainModel2D {
  public:
    TerrainModel2D(std::vector<Site2D> sites, std::vector<Area> areas, Graph graph,
                   std::vector<std::uint32_t> default_outlets, Triangulation triangulation = Triangulation())
        : sites_(std::move(sites)), areas_(std::move(areas)), graph_(std::move(graph)),
          default_outlets_(std::move(default_outlets)), triangulation_(std::move(triangulation)) {}
    std::size_t num() const { return graph_.order(); }  // model.rs:42-44
    const std::vector<Site2D>& sites() const { return sites_; }
    const std::vector<Area>& areas() const { return areas_; }
    const std::vector<std::uint32_t>& default_outlets() const { return default_outlets_; }
    const Graph& graph() const { return graph_; }
    Terrain2D create_terrain_from_result(const std::vector<Elevation>& elevations) const {  // model.rs:62-68
        return Terrain2D(sites_, elevations, TerrainInterpolator2D(sites_, triangulation_));
    }

  private:
    std::vector<Site2D> sites_;
    std::vector<Area> areas_;
    Graph graph_;
    std::vector<std::uint32_t> default_outlets_;
    Triangulation triangulation_;
};

// src/lem/generator.rs:18-26 (+ DeviceError)
enum class GenerationError { InvalidNumberOfParameters, ParametersNotSet, ModelNotSet, DeviceError };

inline const char* to_string(GenerationError e) {
    switch (e) {
        case GenerationError::InvalidNumberOfParameters:
            return "The number of topographical parameters must be equal to the number of sites";
        case GenerationError::ParametersNotSet:
            return "You must set topographical parameters before generating terrain";
        case GenerationError::ModelNotSet: return "You must set `TerrainModel` before generating terrain";
        default: return "CUDA device error (see TerrainGenerator::last_error)";
    }
}

template <class T> class Result {
  public:
    static Result Ok(T v) { Result r; r.value_ = std::move(v); return r; }
    static Result Err(GenerationError e) { Result r; r.error_ = e; return r; }
    bool is_ok() const { return value_.has_value(); }
    bool is_err() const { return !is_ok(); }
    const T& unwrap() const { return value_.value(); }
    GenerationError unwrap_err() const { return error_.value(); }

  private:
    std::optional<T> value_;
    std::optional<GenerationError> error_;
};

// src/lem/generator.rs:36-213
template <class M = TerrainModel2D, class T = Terrain2D> class TerrainGenerator {
  public:
    static TerrainGenerator default_() { return TerrainGenerator(); }
    TerrainGenerator set_model(M model) const { auto g = *this; g.model_ = std::move(model); return g; }
    TerrainGenerator set_parameters(std::vector<TopographicalParameters> p) const {
        auto g = *this; g.parameters_ = std::move(p); return g;
    }
    TerrainGenerator set_max_iteration(Step n) const { auto g = *this; g.max_iteration_ = n; return g; }
    TerrainGenerator set_device(int ordinal) const { auto g = *this; g.device_ = ordinal; return g; }

    Step last_iterations() const { return iterations_; }
    const std::string& last_error() const { return error_; }

    Result<T> generate() {
        if (!model_) return Result<T>::Err(GenerationError::ModelNotSet);  // generator.rs:91-97
        const M& model = *model_;
        const std::size_t num = model.num();
        if (!parameters_) return Result<T>::Err(GenerationError::ParametersNotSet);  // :107-116
        const auto& params = *parameters_;
        if (params.size() != num) return Result<T>::Err(GenerationError::InvalidNumberOfParameters);

        // generator.rs:120-132: explicit outlets ascending, else the model's defaults
        std::vector<std::uint32_t> outlets;
        for (std::size_t i = 0; i < num; ++i)
            if (params[i].is_outlet) outlets.push_back((std::uint32_t)i);
        if (outlets.empty()) outlets = model.default_outlets();

        // flatten AoS -> SoA; tan(max_slope) once per site (generator.rs:194), NaN = None
        std::vector<double> base(num), erod(num), uplift(num), tan_slope;
        bool any_slope = false;
        for (std::size_t i = 0; i < num; ++i) any_slope = any_slope || params[i].max_slope.has_value();
        if (any_slope) tan_slope.assign(num, std::numeric_limits<double>::quiet_NaN());
        for (std::size_t i = 0; i < num; ++i) {
            base[i] = params[i].base_elevation;
            erod[i] = params[i].erodibility;
            uplift[i] = params[i].uplift_rate;
            if (params[i].max_slope) tan_slope[i] = std::tan(*params[i].max_slope);
        }
        // generator.rs:134-138: base + StdRng::seed_from_u64(0).gen::<f64>() * f64::EPSILON
        std::vector<double> initial(num);
        fastlem_host_initial_elevations((std::uint32_t)num, base.data(), initial.data());

        fastlem_ctx* ctx = nullptr;
        if (fastlem_create(&ctx, device_) != FASTLEM_OK) {
            error_ = "fastlem_create failed (no CUDA device? there is no CPU fallback)";
            return Result<T>::Err(GenerationError::DeviceError);
        }
        const Graph& g = model.graph();
        std::vector<double> elevations(num);
        std::uint32_t it = 0;
        int rc = fastlem_set_graph(ctx, (std::uint32_t)num, g.row_ptr.data(), g.col.data(), g.dist.data(),
                                   model.areas().data());
        if (rc == FASTLEM_OK)
            rc = fastlem_set_parameters(ctx, initial.data(), erod.data(), uplift.data(),
                                        any_slope ? tan_slope.data() : nullptr, outlets.data(),
                                        (std::uint32_t)outlets.size());
        if (rc == FASTLEM_OK)
            rc = fastlem_generate(ctx, max_iteration_ ? *max_iteration_ : FASTLEM_UNTIL_STABLE, elevations.data(), &it);
        if (rc != FASTLEM_OK) error_ = fastlem_last_error(ctx);
        fastlem_destroy(ctx);
        if (rc != FASTLEM_OK) return Result<T>::Err(GenerationError::DeviceError);
        iterations_ = it;
        return Result<T>::Ok(model.create_terrain_from_result(elevations));  // generator.rs:212
    }

  private:
    std::optional<M> model_;
    std::optional<std::vector<TopographicalParameters>> parameters_;
    std::optional<Step> max_iteration_;
    int device_ = 0;
    Step iterations_ = 0;
    std::string error_;
};

}  // namespace fastlem
