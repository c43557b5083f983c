// fastlem_oracle.cpp -- TEST INFRASTRUCTURE ONLY (not product code).
//
// Single-threaded CPU restatement of the reference's terrain solve, the
// `TerrainGenerator::generate()` path of TadaTeruki/fastlem 0.1.4:
//   src/lem/generator.rs:90-213      (outer loop, drainage area, response time, elevation)
//   src/lem/stream_tree.rs:72-243    (receivers, root labelling, lake removal)
//   src/lem/drainage_basin.rs:13-46  (per-outlet BFS order)
// plus the parts of un-vendored dependencies the path relies on, restated from
// their published algorithms (none of them is present under /root/reference):
//   rand 0.8.5 StdRng = rand_chacha 0.3 ChaCha12, rand_core 0.6 seed_from_u64 (PCG32),
//   rand 0.8.5 Standard f64 / UniformFloat<f64> sampling,
//   Rust std::collections::BinaryHeap push/pop (sift_up / sift_down_to_bottom),
//   terrain-graph 1.0.1 EdgeAttributedUndirectedGraph (insertion-ordered adjacency).
//
// PARITY UNPINNED: the reference crate cannot be compiled here (no cargo/rustc,
// no network) and its tests hold no golden vectors for this path (tests/*.rs only
// write image.png).  What *is* pinned: the ChaCha block function against the
// public ChaCha8/12/20 zero-key vectors (tests/test_oracle.py).  Everything else
// is a line-by-line restatement, cross-checked by an independent pure-Python
// restatement (oracle/pyref.py).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.  The product (fastlem_b200/) never does.
//
// Build: see oracle/Makefile  (g++ -O3 -ffp-contract=off, no fast-math).

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// rand_chacha 0.3 ChaCha{8,12,20}Rng core: "expand 32-byte k", 8 key words, 64-bit block
// counter in words 12-13, 64-bit stream id in words 14-15 (0).  [dep: rand_chacha 0.3.1]
// ---------------------------------------------------------------------------------------------
inline uint32_t rotl32(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }

inline void quarter(uint32_t* x, int a, int b, int c, int d) {
    x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl32(x[d], 16);
    x[c] += x[d]; x[b] ^= x[c]; x[b] = rotl32(x[b], 12);
    x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl32(x[d], 8);
    x[c] += x[d]; x[b] ^= x[c]; x[b] = rotl32(x[b], 7);
}

void chacha_block(const uint32_t key[8], uint64_t counter, int double_rounds, uint32_t out[16]) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                      key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                      (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
    uint32_t x[16];
    std::memcpy(x, s, sizeof(x));
    for (int r = 0; r < double_rounds; ++r) {
        quarter(x, 0, 4, 8, 12); quarter(x, 1, 5, 9, 13); quarter(x, 2, 6, 10, 14); quarter(x, 3, 7, 11, 15);
        quarter(x, 0, 5, 10, 15); quarter(x, 1, 6, 11, 12); quarter(x, 2, 7, 8, 13); quarter(x, 3, 4, 9, 14);
    }
    for (int i = 0; i < 16; ++i) out[i] = x[i] + s[i];
}

// StdRng (rand 0.8.5) = ChaCha12Rng; BlockRng hands out the words of consecutive blocks in order.
struct StdRng {
    uint32_t key[8];
    uint64_t counter = 0;
    uint32_t buf[16];
    int idx = 16;
    uint32_t next_u32() {
        if (idx >= 16) { chacha_block(key, counter++, 6, buf); idx = 0; }
        return buf[idx++];
    }
    // BlockRng::next_u64: two consecutive u32 words, low word first.
    uint64_t next_u64() {
        uint64_t lo = next_u32();
        uint64_t hi = next_u32();
        return lo | (hi << 32);
    }
    // rand 0.8.5 `Standard` for f64: 53 high bits * 2^-53.
    double gen_f64() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
    // rand 0.8.5 UniformFloat<f64>::sample_single (gen_range(low..high)).
    double gen_range(double low, double high) {
        double scale = high - low;
        for (;;) {
            uint64_t bits = (next_u64() >> 12) | (0x3FFull << 52);
            double v12; std::memcpy(&v12, &bits, 8);
            double res = (v12 - 1.0) * scale + low;
            if (res < high) return res;
        }
    }
    static StdRng from_seed(const uint8_t seed[32]) {
        StdRng r;
        for (int i = 0; i < 8; ++i)
            r.key[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) |
                       ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
        return r;
    }
    // rand_core 0.6 SeedableRng::seed_from_u64: PCG32 stream fills the seed 4 bytes at a time.
    static void seed_bytes_from_u64(uint64_t state, uint8_t seed[32]) {
        const uint64_t MUL = 6364136223846793005ull, INC = 11634580027462260723ull;
        for (int i = 0; i < 8; ++i) {
            state = state * MUL + INC;
            uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
            uint32_t rot = (uint32_t)(state >> 59);
            uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
            seed[4 * i] = (uint8_t)x; seed[4 * i + 1] = (uint8_t)(x >> 8);
            seed[4 * i + 2] = (uint8_t)(x >> 16); seed[4 * i + 3] = (uint8_t)(x >> 24);
        }
    }
    static StdRng seed_from_u64(uint64_t s) {
        uint8_t seed[32];
        seed_bytes_from_u64(s, seed);
        return from_seed(seed);
    }
};

// ---------------------------------------------------------------------------------------------
// terrain-graph 1.0.1 EdgeAttributedUndirectedGraph<f64>, viewed through a CSR whose rows are
// the insertion-ordered adjacency lists (`neighbors_of(i)` = row i in order).
// ---------------------------------------------------------------------------------------------
struct Graph {
    uint32_t n;
    const uint32_t* row_ptr;
    const uint32_t* col;
    const double* dist;
    // has_edge(a,b): first match in a's list -> (true, attr) else (false, default)
    bool has_edge(uint32_t a, uint32_t b, double* attr) const {
        for (uint32_t s = row_ptr[a]; s < row_ptr[a + 1]; ++s)
            if (col[s] == b) { *attr = dist[s]; return true; }
        *attr = 0.0;
        return false;
    }
};

// ---------------------------------------------------------------------------------------------
// Rust std BinaryHeap<RidgeElement> (max-heap); RidgeElement ordering from
// src/lem/stream_tree.rs:15-44: cmp(a,b) = b.dist.partial_cmp(a.dist), so `a <= b` <=> a.dist >= b.dist.
// ---------------------------------------------------------------------------------------------
struct RidgeElement { uint32_t index; double dist; };
inline bool le(const RidgeElement& a, const RidgeElement& b) { return a.dist >= b.dist; }

struct BinaryHeap {
    std::vector<RidgeElement> data;
    void sift_up(size_t start, size_t pos) {
        RidgeElement elt = data[pos];
        while (pos > start) {
            size_t parent = (pos - 1) / 2;
            if (le(elt, data[parent])) break;
            data[pos] = data[parent];
            pos = parent;
        }
        data[pos] = elt;
    }
    void sift_down_to_bottom(size_t pos) {
        size_t end = data.size();
        size_t start = pos;
        RidgeElement elt = data[pos];
        size_t child = 2 * pos + 1;
        size_t lim = end >= 2 ? end - 2 : 0;  // end.saturating_sub(2)
        while (child <= lim) {
            child += le(data[child], data[child + 1]) ? 1 : 0;
            data[pos] = data[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1) {
            data[pos] = data[child];
            pos = child;
        }
        data[pos] = elt;
        sift_up(start, pos);
    }
    void push(RidgeElement e) {
        size_t old_len = data.size();
        data.push_back(e);
        sift_up(0, old_len);
    }
    bool pop(RidgeElement* out) {
        if (data.empty()) return false;
        RidgeElement item = data.back();
        data.pop_back();
        if (!data.empty()) {
            std::swap(item, data[0]);
            sift_down_to_bottom(0);
        }
        *out = item;
        return true;
    }
};

const uint32_t NONE = 0xFFFFFFFFu;

// stream_tree.rs:101-107
std::vector<uint8_t> create_outlet_table(uint32_t num, const uint32_t* outlets, uint32_t n_outlets) {
    std::vector<uint8_t> is_outlet(num, 0);
    for (uint32_t k = 0; k < n_outlets; ++k) is_outlet[outlets[k]] = 1;
    return is_outlet;
}

// stream_tree.rs:109-137
std::vector<uint32_t> construct_initial_stream_tree(uint32_t num, const double* elevations, const Graph& g,
                                                    const std::vector<uint8_t>& is_outlet) {
    std::vector<uint32_t> next(num);
    for (uint32_t i = 0; i < num; ++i) next[i] = i;
    for (uint32_t i = 0; i < num; ++i) {
        if (is_outlet[i]) continue;
        double steepest_slope = 0.0;
        for (uint32_t s = g.row_ptr[i]; s < g.row_ptr[i + 1]; ++s) {
            uint32_t j = g.col[s];
            if (elevations[i] > elevations[j]) {
                double distance = g.dist[s];
                double down_hill_slope = (elevations[i] - elevations[j]) / distance;
                if (down_hill_slope > steepest_slope) {
                    steepest_slope = down_hill_slope;
                    next[i] = j;
                }
            }
        }
    }
    return next;
}

// stream_tree.rs:139-173
bool find_roots_with_lakes(uint32_t num, const std::vector<uint8_t>& is_outlet, const std::vector<uint32_t>& next,
                           std::vector<uint32_t>& subroot) {
    subroot.assign(num, NONE);
    for (uint32_t i = 0; i < num; ++i) if (is_outlet[i]) subroot[i] = i;
    bool has_lake = false;
    for (uint32_t i = 0; i < num; ++i) {
        if (subroot[i] != NONE) continue;
        uint32_t iv = i;
        while (subroot[iv] == NONE && iv != next[iv]) iv = next[iv];
        uint32_t ir;
        if (subroot[iv] == NONE) { has_lake = true; ir = iv; } else { ir = subroot[iv]; }
        iv = i;
        while (subroot[iv] == NONE && iv != next[iv]) { subroot[iv] = ir; iv = next[iv]; }
        subroot[iv] = ir;
    }
    return has_lake;
}

// stream_tree.rs:175-243.  `pop_order` (optional) receives the sequence number of each node's first pop.
std::vector<uint32_t> remove_lakes_from_stream_tree(const std::vector<uint32_t>& next_in, uint32_t num, const Graph& g,
                                                    const uint32_t* outlets, uint32_t n_outlets,
                                                    const std::vector<uint32_t>& subroot, uint32_t* pop_order) {
    std::vector<uint32_t> root(num, NONE);
    BinaryHeap ridgestack;
    ridgestack.data.reserve(num);
    for (uint32_t k = 0; k < n_outlets; ++k) {
        uint32_t outlet = outlets[k];
        root[outlet] = outlet;
        ridgestack.push(RidgeElement{outlet, 0.0});
    }
    std::vector<uint8_t> visited(num, 0);
    std::vector<uint32_t> next = next_in;
    uint32_t seq = 0;
    RidgeElement element;
    while (ridgestack.pop(&element)) {
        uint32_t i = element.index;
        if (visited[i]) continue;
        if (pop_order) pop_order[i] = seq;
        ++seq;
        for (uint32_t s = g.row_ptr[i]; s < g.row_ptr[i + 1]; ++s) {
            uint32_t j = g.col[s];
            if (visited[j]) continue;
            if (root[subroot[j]] == NONE) {
                uint32_t k = j;
                uint32_t nk = i;
                for (;;) {
                    if (next[k] != k) {
                        uint32_t tmp = next[k];   // flip flow
                        next[k] = nk;
                        nk = k;
                        k = tmp;
                    } else {
                        break;
                    }
                }
                next[k] = nk;
                root[subroot[j]] = root[subroot[i]];
            }
            ridgestack.push(RidgeElement{j, g.dist[s]});
        }
        root[i] = root[subroot[i]];
        visited[i] = 1;
    }
    return next;
}

// stream_tree.rs:72-99
std::vector<uint32_t> stream_tree_construct(uint32_t num, const double* elevations, const Graph& g,
                                            const uint32_t* outlets, uint32_t n_outlets,
                                            std::vector<uint32_t>* next_initial, std::vector<uint32_t>* subroot_out,
                                            bool* has_lake_out) {
    std::vector<uint8_t> is_outlet = create_outlet_table(num, outlets, n_outlets);
    std::vector<uint32_t> next = construct_initial_stream_tree(num, elevations, g, is_outlet);
    std::vector<uint32_t> subroot;
    bool has_lake = find_roots_with_lakes(num, is_outlet, next, subroot);
    if (next_initial) *next_initial = next;
    if (subroot_out) *subroot_out = subroot;
    if (has_lake_out) *has_lake_out = has_lake;
    if (!has_lake) return next;
    return remove_lakes_from_stream_tree(next, num, g, outlets, n_outlets, subroot, nullptr);
}

// drainage_basin.rs:13-36
void drainage_basin_construct(uint32_t outlet, const std::vector<uint32_t>& next, const Graph& g,
                              std::vector<uint32_t>& traversal) {
    traversal.clear();
    traversal.push_back(outlet);
    size_t i = 0;
    for (;;) {
        uint32_t it = traversal[i];
        for (uint32_t s = g.row_ptr[it]; s < g.row_ptr[it + 1]; ++s) {
            uint32_t jt = g.col[s];
            if (next[jt] == it) traversal.push_back(jt);
        }
        i += 1;
        if (i >= traversal.size()) break;
    }
}

const double DEFAULT_M_EXP = 0.5;  // generator.rs:16

// One pass of the loop body, generator.rs:141-205.  Returns `changed`.
// max_slope: radians per node, NaN = None (may be null = all None).
bool iterate_once(const Graph& g, uint32_t num, const double* areas, const double* erodibility,
                  const double* uplift_rate, const double* max_slope, const uint32_t* outlets, uint32_t n_outlets,
                  double* elevations, uint32_t* next_out, uint32_t* next_initial_out, uint32_t* subroot_out,
                  int* has_lake_out, double* drainage_out, double* response_out, uint32_t* order_out) {
    std::vector<uint32_t> next_initial, subroot;
    bool has_lake = false;
    std::vector<uint32_t> next = stream_tree_construct(num, elevations, g, outlets, n_outlets, &next_initial,
                                                       &subroot, &has_lake);
    std::vector<double> drainage_areas(areas, areas + num);
    std::vector<double> response_times(num, 0.0);
    bool changed = false;
    const double m_exp = DEFAULT_M_EXP;
    std::vector<uint32_t> traversal;
    uint32_t order_seq = 0;
    if (order_out) for (uint32_t i = 0; i < num; ++i) order_out[i] = NONE;

    for (uint32_t k = 0; k < n_outlets; ++k) {
        uint32_t outlet = outlets[k];
        drainage_basin_construct(outlet, next, g, traversal);
        if (order_out) for (uint32_t i : traversal) order_out[i] = order_seq++;

        // generator.rs:154-159 (downstream = reverse traversal)
        for (size_t t = traversal.size(); t-- > 0;) {
            uint32_t i = traversal[t];
            uint32_t j = next[i];
            if (j != i) drainage_areas[j] += drainage_areas[i];
        }
        // generator.rs:162-174
        for (uint32_t i : traversal) {
            uint32_t j = next[i];
            double edge;
            double distance = g.has_edge(i, j, &edge) ? edge : 1.0;
            // powf(0.5): release-mode LLVM folds llvm.pow(x, 0.5) into sqrt(x) ("FP discipline")
            (void)m_exp;
            double celerity = erodibility[i] * std::sqrt(drainage_areas[i]);
            response_times[i] += response_times[j] + 1.0 / celerity * distance;
        }
        // generator.rs:177-203
        for (uint32_t i : traversal) {
            double new_elevation =
                elevations[outlet] + uplift_rate[i] * std::fmax(response_times[i] - response_times[outlet], 0.0);
            if (max_slope && !std::isnan(max_slope[i])) {
                uint32_t j = next[i];
                double edge;
                double distance = g.has_edge(i, j, &edge) ? edge : 1.0;
                double ms = std::tan(max_slope[i]);
                double slope = (new_elevation - elevations[j]) / distance;
                if (slope > ms) new_elevation = elevations[j] + ms * distance;
            }
            changed |= new_elevation != elevations[i];
            elevations[i] = new_elevation;
        }
    }
    if (next_out) std::memcpy(next_out, next.data(), sizeof(uint32_t) * num);
    if (next_initial_out) std::memcpy(next_initial_out, next_initial.data(), sizeof(uint32_t) * num);
    if (subroot_out) std::memcpy(subroot_out, subroot.data(), sizeof(uint32_t) * num);
    if (has_lake_out) *has_lake_out = has_lake ? 1 : 0;
    if (drainage_out) std::memcpy(drainage_out, drainage_areas.data(), sizeof(double) * num);
    if (response_out) std::memcpy(response_out, response_times.data(), sizeof(double) * num);
    return changed;
}

}  // namespace

extern "C" {

// KAT hook: one ChaCha block with `double_rounds` (4/6/10 = ChaCha8/12/20).
void fo_chacha_block(const uint32_t key[8], uint64_t counter, int double_rounds, uint32_t out[16]) {
    chacha_block(key, counter, double_rounds, out);
}

void fo_seed_from_u64(uint64_t state, uint8_t seed_out[32]) { StdRng::seed_bytes_from_u64(state, seed_out); }

// generator.rs:134-138: elevations[i] = base[i] + rng.gen::<f64>() * f64::EPSILON, rng = seed_from_u64(0).
void fo_initial_elevations(uint32_t n, const double* base_elevation, double* out) {
    StdRng rng = StdRng::seed_from_u64(0);
    for (uint32_t i = 0; i < n; ++i)
        out[i] = base_elevation[i] + rng.gen_f64() * std::numeric_limits<double>::epsilon();
}

// first `n` gen::<f64>() values of seed_from_u64(seed) (test hook)
void fo_gen_f64(uint64_t seed, uint32_t n, double* out) {
    StdRng rng = StdRng::seed_from_u64(seed);
    for (uint32_t i = 0; i < n; ++i) out[i] = rng.gen_f64();
}

// builder.rs:38-46: from_random_sites uses StdRng::from_seed([0u8;32]) and gen_range per coordinate (x then y).
// `seed_byte` fills the 32-byte seed (0 reproduces the reference); xy_out is interleaved x,y.
void fo_random_sites(uint32_t n, uint8_t seed_byte, double min_x, double min_y, double max_x, double max_y,
                     double* xy_out) {
    uint8_t seed[32];
    std::memset(seed, seed_byte, 32);
    StdRng rng = StdRng::from_seed(seed);
    for (uint32_t i = 0; i < n; ++i) {
        xy_out[2 * i] = rng.gen_range(min_x, max_x);
        xy_out[2 * i + 1] = rng.gen_range(min_y, max_y);
    }
}

// Pop order of the heap flood of stream_tree.rs:175-243 (first pop of each node; NONE if never popped).
void fo_flood_order(uint32_t n, const uint32_t* row_ptr, const uint32_t* col, const double* dist,
                    const uint32_t* outlets, uint32_t n_outlets, uint32_t* order_out) {
    Graph g{n, row_ptr, col, dist};
    std::vector<uint32_t> next(n), subroot(n);
    // a lake-free dummy forest: every node its own outlet-rooted label is not needed for the order itself,
    // the order depends only on (graph, outlets); use next[i]=i / subroot[i]=first outlet so no path is flipped.
    for (uint32_t i = 0; i < n; ++i) { next[i] = i; subroot[i] = n_outlets ? outlets[0] : 0; }
    for (uint32_t i = 0; i < n; ++i) order_out[i] = NONE;
    if (n_outlets == 0) return;
    remove_lakes_from_stream_tree(next, n, g, outlets, n_outlets, subroot, order_out);
}

// StreamTree::construct with all intermediates.
void fo_stream_tree(uint32_t n, const uint32_t* row_ptr, const uint32_t* col, const double* dist,
                    const double* elevations, const uint32_t* outlets, uint32_t n_outlets, uint32_t* next_out,
                    uint32_t* next_initial_out, uint32_t* subroot_out, int* has_lake_out) {
    Graph g{n, row_ptr, col, dist};
    std::vector<uint32_t> ni, sr;
    bool hl = false;
    std::vector<uint32_t> next = stream_tree_construct(n, elevations, g, outlets, n_outlets, &ni, &sr, &hl);
    std::memcpy(next_out, next.data(), sizeof(uint32_t) * n);
    if (next_initial_out) std::memcpy(next_initial_out, ni.data(), sizeof(uint32_t) * n);
    if (subroot_out) std::memcpy(subroot_out, sr.data(), sizeof(uint32_t) * n);
    if (has_lake_out) *has_lake_out = hl ? 1 : 0;
}

// One loop body (generator.rs:141-205) in place on `elevations`; returns changed (0/1).  Any output may be null.
int fo_iterate_once(uint32_t n, const uint32_t* row_ptr, const uint32_t* col, const double* dist, const double* areas,
                    const double* erodibility, const double* uplift_rate, const double* max_slope,
                    const uint32_t* outlets, uint32_t n_outlets, double* elevations, uint32_t* next_out,
                    uint32_t* next_initial_out, uint32_t* subroot_out, int* has_lake_out, double* drainage_out,
                    double* response_out, uint32_t* order_out) {
    Graph g{n, row_ptr, col, dist};
    return iterate_once(g, n, areas, erodibility, uplift_rate, max_slope, outlets, n_outlets, elevations, next_out,
                        next_initial_out, subroot_out, has_lake_out, drainage_out, response_out, order_out)
               ? 1 : 0;
}

// generator.rs:140-210: iterate from `elevations` (in/out; already base+noise) until stable or max_iteration.
// max_iteration = UINT32_MAX means "not set".  Returns the number of loop bodies executed.
uint32_t fo_generate(uint32_t n, const uint32_t* row_ptr, const uint32_t* col, const double* dist, const double* areas,
                     const double* erodibility, const double* uplift_rate, const double* max_slope,
                     const uint32_t* outlets, uint32_t n_outlets, uint32_t max_iteration, double* elevations) {
    Graph g{n, row_ptr, col, dist};
    uint32_t it = 0;
    for (; it < max_iteration; ) {
        bool changed = iterate_once(g, n, areas, erodibility, uplift_rate, max_slope, outlets, n_outlets, elevations,
                                    nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        ++it;
        if (!changed) break;
    }
    return it;
}

}  // extern "C"
