// landscape_evolution.cpp -- the reference's examples/landscape_evolution.rs:18-62 through the C++ mirror
// (fastlem_b200/host/fastlem.hpp).  The model (sites + Delaunay graph) is read from a small binary file written
// by tools/workloads.py, because the graph build is outside this path.
//
//   ./landscape_evolution model.bin elevations.bin [max_slope_radians] [max_iteration] [image_size image.bin]
//   exit status 0 = ok, 2..5 = GenerationError variant
// With image_size > 0 the render loop of the example (:36-62) runs too: get_elevation per pixel, bound_max = 100,
// written as image_size^2 doubles (row = imgy, NaN where the reference skips the pixel); the one-call raster must
// give the same image (exit status 13 otherwise).
// file format (little endian): u32 n, u32 nnz, u32 n_outlets, u32 row_ptr[n+1], u32 col[nnz], f64 dist[nnz],
//                              f64 areas[n], u32 default_outlets[n_outlets], f64 sites_xy[2n],
//                              optionally u32 n_triangles, u32 triangles[3T], u32 halfedges[3T]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../fastlem_b200/host/fastlem.hpp"

template <class T> static bool rd(FILE* f, std::vector<T>& v, size_t n) {
    v.resize(n);
    return n == 0 || std::fread(v.data(), sizeof(T), n, f) == n;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s model.bin out.bin [max_slope] [max_iteration]\n", argv[0]); return 1; }
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) { std::perror("model"); return 1; }
    uint32_t hdr[3];
    if (std::fread(hdr, 4, 3, f) != 3) return 1;
    const uint32_t n = hdr[0], nnz = hdr[1], n_out = hdr[2];
    fastlem::Graph g;
    std::vector<double> areas, xy;
    std::vector<uint32_t> outlets;
    if (!rd(f, g.row_ptr, (size_t)n + 1) || !rd(f, g.col, nnz) || !rd(f, g.dist, nnz) || !rd(f, areas, n) ||
        !rd(f, outlets, n_out) || !rd(f, xy, (size_t)2 * n)) { std::fprintf(stderr, "short model file\n"); return 1; }
    fastlem::Triangulation tri;
    uint32_t n_tri = 0;
    if (std::fread(&n_tri, 4, 1, f) == 1 &&
        (!rd(f, tri.triangles, (size_t)3 * n_tri) || !rd(f, tri.halfedges, (size_t)3 * n_tri))) {
        std::fprintf(stderr, "short triangulation\n");
        return 1;
    }
    std::fclose(f);
    std::vector<fastlem::Site2D> sites(n);
    for (uint32_t i = 0; i < n; ++i) sites[i] = fastlem::Site2D{xy[2 * i], xy[2 * i + 1]};
    fastlem::TerrainModel2D model(sites, areas, g, outlets, tri);

    auto proto = fastlem::TopographicalParameters::default_().set_erodibility(1.0);
    if (argc > 3 && std::atof(argv[3]) > 0.0) proto = proto.set_max_slope(std::atof(argv[3]));
    std::vector<fastlem::TopographicalParameters> params(n, proto);

    auto gen = fastlem::TerrainGenerator<>::default_().set_model(model).set_parameters(params);
    if (argc > 4) gen = gen.set_max_iteration((fastlem::Step)std::atoi(argv[4]));
    auto terrain = gen.generate();
    if (terrain.is_err()) {
        std::fprintf(stderr, "generate: %s %s\n", fastlem::to_string(terrain.unwrap_err()), gen.last_error().c_str());
        return 2 + (int)terrain.unwrap_err();
    }
    FILE* o = std::fopen(argv[2], "wb");
    const auto& e = terrain.unwrap().elevations();
    std::fwrite(e.data(), sizeof(double), e.size(), o);
    std::fclose(o);
    std::printf("iterations %u\n", gen.last_iterations());

    // examples/landscape_evolution.rs:36-62
    const uint32_t img = argc > 6 ? (uint32_t)std::atoi(argv[5]) : 0u;
    if (img > 0) {
        const fastlem::Terrain2D& t = terrain.unwrap();
        const fastlem::Site2D bound_max{100.0, 100.0};
        const double nan = std::nan("");
        std::vector<double> image((size_t)img * img, nan);
        for (uint32_t imgx = 0; imgx < img; ++imgx)
            for (uint32_t imgy = 0; imgy < img; ++imgy) {
                const double x = bound_max.x * ((double)imgx / (double)img);
                const double y = bound_max.y * ((double)imgy / (double)img);
                const auto elevation = t.get_elevation(fastlem::Site2D{x, y});
                if (elevation) image[(size_t)imgy * img + imgx] = *elevation;
            }
        const fastlem_raster r{0.0, 0.0, bound_max.x, bound_max.y, 0.0, img, img, 0u, img};
        const std::vector<double> whole = t.raster(r);
        if (std::memcmp(whole.data(), image.data(), sizeof(double) * image.size()) != 0) {
            for (size_t i = 0; i < image.size(); ++i)
                if (!(whole[i] == image[i]) && !(whole[i] != whole[i] && image[i] != image[i])) return 13;
        }
        FILE* io = std::fopen(argv[6], "wb");
        std::fwrite(image.data(), sizeof(double), image.size(), io);
        std::fclose(io);
    }

    // the three validation errors of generator.rs:91-116
    auto e1 = fastlem::TerrainGenerator<>::default_().generate();
    auto e2 = fastlem::TerrainGenerator<>::default_().set_model(model).generate();
    auto e3 = fastlem::TerrainGenerator<>::default_().set_model(model)
                  .set_parameters(std::vector<fastlem::TopographicalParameters>(3)).generate();
    if (!(e1.is_err() && e1.unwrap_err() == fastlem::GenerationError::ModelNotSet)) return 10;
    if (!(e2.is_err() && e2.unwrap_err() == fastlem::GenerationError::ParametersNotSet)) return 11;
    if (!(e3.is_err() && e3.unwrap_err() == fastlem::GenerationError::InvalidNumberOfParameters)) return 12;
    return 0;
}
