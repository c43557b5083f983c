// fl_host.cpp -- host-side pieces of the generate() prologue that the C++/Python host mirrors need.
//
// In the Rust integration the shim keeps calling the `rand` crate for the epsilon noise
// (reference src/lem/generator.rs:134-138), so this file is NOT on the Rust path; it exists so the
// C++ and Python mirrors of TerrainGenerator can produce the same initial elevations without Rust:
//   rng = StdRng::seed_from_u64(0); elevations[i] = base[i] + rng.gen::<f64>() * f64::EPSILON
// StdRng (rand 0.8.5) is ChaCha with 12 rounds; seed_from_u64 (rand_core 0.6) expands the u64 with
// a PCG32 generator; gen::<f64>() keeps the top 53 bits of next_u64().
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/fastlem_b200.h"

namespace {

struct ChaCha12 {
    uint32_t st[16];
    uint32_t out[16];
    int used = 16;

    explicit ChaCha12(const uint32_t key[8]) {
        static const uint32_t sigma[4] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
        std::memcpy(st, sigma, 16);
        std::memcpy(st + 4, key, 32);
        st[12] = st[13] = st[14] = st[15] = 0;  // 64-bit block counter, 64-bit stream id
    }

    static inline uint32_t rol(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

    void refill() {
        uint32_t w[16];
        std::memcpy(w, st, 64);
#define FL_QR(a, b, c, d)                          \
    w[a] += w[b]; w[d] = rol(w[d] ^ w[a], 16);     \
    w[c] += w[d]; w[b] = rol(w[b] ^ w[c], 12);     \
    w[a] += w[b]; w[d] = rol(w[d] ^ w[a], 8);      \
    w[c] += w[d]; w[b] = rol(w[b] ^ w[c], 7);
        for (int pair = 0; pair < 6; ++pair) {  // 12 rounds = 6 column+diagonal pairs
            FL_QR(0, 4, 8, 12) FL_QR(1, 5, 9, 13) FL_QR(2, 6, 10, 14) FL_QR(3, 7, 11, 15)
            FL_QR(0, 5, 10, 15) FL_QR(1, 6, 11, 12) FL_QR(2, 7, 8, 13) FL_QR(3, 4, 9, 14)
        }
#undef FL_QR
        for (int k = 0; k < 16; ++k) out[k] = w[k] + st[k];
        if (++st[12] == 0) ++st[13];
        used = 0;
    }

    uint32_t word() {
        if (used == 16) refill();
        return out[used++];
    }

    uint64_t next_u64() {
        uint64_t lo = word();
        return lo | ((uint64_t)word() << 32);
    }
};

void pcg32_expand(uint64_t state, uint32_t key[8]) {
    for (int k = 0; k < 8; ++k) {
        state = state * 6364136223846793005ull + 11634580027462260723ull;
        uint32_t x = (uint32_t)(((state >> 18) ^ state) >> 27);
        uint32_t r = (uint32_t)(state >> 59);
        key[k] = (x >> r) | (x << ((32u - r) & 31u));  // little-endian bytes == the key word itself
    }
}

}  // namespace

extern "C" void fastlem_host_initial_elevations(uint32_t n, const double* base_elevation, double* out) {
    uint32_t key[8];
    pcg32_expand(0, key);
    ChaCha12 rng(key);
    const double eps = 2.220446049250313e-16;      // f64::EPSILON
    const double scale = 1.0 / 9007199254740992.0;  // 2^-53
    for (uint32_t i = 0; i < n; ++i) {
        double u = (double)(rng.next_u64() >> 11) * scale;
        out[i] = base_elevation[i] + u * eps;
    }
}

// generator.rs:194: `max_slope.tan()` -- Rust's f64::tan is libm's tan
extern "C" void fastlem_host_tan_max_slope(uint32_t n, const double* max_slope, double* out) {
    for (uint32_t i = 0; i < n; ++i) out[i] = std::tan(max_slope[i]);
}

// The graph build of TerrainModel2DBulider::build (reference src/models/surface/builder.rs:252-268), straight into the
// boundary format of fastlem_set_graph: for every triangle (a, b, c) of the builder's triangulation, in order, the
// half-edges a->b, b->c, c->a with from < to become edges; add_edge(u, v, w) appends (v, w) to u's list and (u, w) to
// v's list (terrain-graph), w = Site2D::distance (sites.rs:27-29).  Row i of the CSR is therefore
// graph.neighbors_of(i) in iteration order, without building the Vec<Vec<..>> graph and walking it again.
extern "C" int fastlem_host_graph_from_triangles(uint32_t n_sites, const double* sites_xy, uint32_t n_triangles,
                                                 const uint32_t* triangles, uint32_t* row_ptr, uint32_t* col,
                                                 double* dist, uint64_t capacity, uint64_t* nnz_out) {
    if ((n_sites && !sites_xy) || (n_triangles && !triangles) || !row_ptr) return FASTLEM_E_INVALID;
    // pass 1: degrees
    for (uint32_t i = 0; i <= n_sites; ++i) row_ptr[i] = 0;
    uint64_t nnz = 0;
    for (uint32_t t = 0; t < n_triangles; ++t) {
        const uint32_t* v = triangles + 3 * (size_t)t;
        for (int k = 0; k < 3; ++k) {
            const uint32_t a = v[k], b = v[(k + 1) % 3];
            if (a >= n_sites || b >= n_sites) return FASTLEM_E_INVALID;
            if (a < b) { ++row_ptr[a + 1]; ++row_ptr[b + 1]; nnz += 2; }
        }
    }
    if (nnz >= 0xFFFFFFFFull) return FASTLEM_E_INVALID;  // CSR offsets are uint32
    if (nnz_out) *nnz_out = nnz;
    for (uint32_t i = 0; i < n_sites; ++i) row_ptr[i + 1] += row_ptr[i];
    if (!col && !dist) return FASTLEM_OK;  // size query
    if (!col || !dist || capacity < nnz) return FASTLEM_E_INVALID;
    // pass 2: fill in insertion order
    std::vector<uint32_t> cursor(row_ptr, row_ptr + n_sites);
    for (uint32_t t = 0; t < n_triangles; ++t) {
        const uint32_t* v = triangles + 3 * (size_t)t;
        for (int k = 0; k < 3; ++k) {
            const uint32_t a = v[k], b = v[(k + 1) % 3];
            if (!(a < b)) continue;
            const double dx = sites_xy[2 * (size_t)a] - sites_xy[2 * (size_t)b];
            const double dy = sites_xy[2 * (size_t)a + 1] - sites_xy[2 * (size_t)b + 1];
            const double w = std::sqrt(dx * dx + dy * dy);
            col[cursor[a]] = b; dist[cursor[a]++] = w;
            col[cursor[b]] = a; dist[cursor[b]++] = w;
        }
    }
    return FASTLEM_OK;
}
