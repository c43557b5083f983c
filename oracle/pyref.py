"""pyref.py -- second, independent restatement of the reference path in pure Python (TEST INFRASTRUCTURE ONLY).

Written straight from the Rust sources, separately from fastlem_oracle.cpp, and deliberately naive: adjacency
is a list of lists of (index, length) like terrain-graph's Vec<Vec<(usize, f64)>>, the heap is a small class
that replays std::collections::BinaryHeap, Python floats are IEEE doubles.  Small cases only.  It is used to
(1) cross-check the C++ oracle and (2) generate the golden vectors under tests/golden/ (tools/make_golden.py).

PARITY UNPINNED (same caveat as fastlem_oracle.cpp): the Rust crate cannot be run here.

References: src/lem/generator.rs:118-210, src/lem/stream_tree.rs:72-243, src/lem/drainage_basin.rs:13-46.
"""
import math
import struct

MASK32 = 0xFFFFFFFF
MASK64 = 0xFFFFFFFFFFFFFFFF


# ---- rand 0.8.5 StdRng (ChaCha12) -------------------------------------------------------------------
def _rotl(v, c):
    return ((v << c) & MASK32) | (v >> (32 - c))


def _qr(x, a, b, c, d):
    x[a] = (x[a] + x[b]) & MASK32; x[d] = _rotl(x[d] ^ x[a], 16)
    x[c] = (x[c] + x[d]) & MASK32; x[b] = _rotl(x[b] ^ x[c], 12)
    x[a] = (x[a] + x[b]) & MASK32; x[d] = _rotl(x[d] ^ x[a], 8)
    x[c] = (x[c] + x[d]) & MASK32; x[b] = _rotl(x[b] ^ x[c], 7)


def chacha_block(key_words, counter, double_rounds):
    init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + \
           [counter & MASK32, (counter >> 32) & MASK32, 0, 0]
    x = list(init)
    for _ in range(double_rounds):
        _qr(x, 0, 4, 8, 12); _qr(x, 1, 5, 9, 13); _qr(x, 2, 6, 10, 14); _qr(x, 3, 7, 11, 15)
        _qr(x, 0, 5, 10, 15); _qr(x, 1, 6, 11, 12); _qr(x, 2, 7, 8, 13); _qr(x, 3, 4, 9, 14)
    return [(x[i] + init[i]) & MASK32 for i in range(16)]


class StdRng:
    def __init__(self, seed32):
        self.key = list(struct.unpack("<8I", bytes(seed32)))
        self.counter = 0
        self.buf = []

    @classmethod
    def seed_from_u64(cls, state):
        seed = b""
        for _ in range(8):
            state = (state * 6364136223846793005 + 11634580027462260723) & MASK64
            xorshifted = (((state >> 18) ^ state) >> 27) & MASK32
            rot = state >> 59
            x = ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & MASK32
            seed += struct.pack("<I", x)
        return cls(seed)

    def next_u32(self):
        if not self.buf:
            self.buf = chacha_block(self.key, self.counter, 6)
            self.counter += 1
        return self.buf.pop(0)

    def next_u64(self):
        lo = self.next_u32()
        hi = self.next_u32()
        return lo | (hi << 32)

    def gen_f64(self):
        return (self.next_u64() >> 11) * (1.0 / (1 << 53))


def initial_elevations(base):
    rng = StdRng.seed_from_u64(0)
    eps = 2.0 ** -52
    return [b + rng.gen_f64() * eps for b in base]


# ---- std::collections::BinaryHeap<RidgeElement> -----------------------------------------------------
class RidgeHeap:
    """Max-heap under Ord for RidgeElement (stream_tree.rs:34-38): a <= b  <=>  a.dist >= b.dist."""

    def __init__(self):
        self.data = []

    @staticmethod
    def _le(a, b):
        return a[1] >= b[1]

    def _sift_up(self, start, pos):
        elt = self.data[pos]
        while pos > start:
            parent = (pos - 1) // 2
            if self._le(elt, self.data[parent]):
                break
            self.data[pos] = self.data[parent]
            pos = parent
        self.data[pos] = elt

    def push(self, item):
        self.data.append(item)
        self._sift_up(0, len(self.data) - 1)

    def pop(self):
        if not self.data:
            return None
        item = self.data.pop()
        if self.data:
            item, self.data[0] = self.data[0], item
            end = len(self.data)
            pos = 0
            elt = self.data[0]
            child = 1
            while child <= max(end - 2, 0) and end >= 2:
                if self._le(self.data[child], self.data[child + 1]):
                    child += 1
                self.data[pos] = self.data[child]
                pos = child
                child = 2 * pos + 1
            if child == end - 1:
                self.data[pos] = self.data[child]
                pos = child
            self.data[pos] = elt
            self._sift_up(0, pos)
        return item


# ---- graph ------------------------------------------------------------------------------------------
def adjacency_from_csr(row_ptr, col, dist):
    return [[(int(col[s]), float(dist[s])) for s in range(int(row_ptr[i]), int(row_ptr[i + 1]))]
            for i in range(len(row_ptr) - 1)]


def has_edge(adj, a, b):
    for (j, w) in adj[a]:
        if j == b:
            return True, w
    return False, 0.0


# ---- stream_tree.rs ---------------------------------------------------------------------------------
def stream_tree(adj, elevations, outlets):
    num = len(adj)
    is_outlet = [False] * num
    for o in outlets:
        is_outlet[o] = True
    nxt = list(range(num))
    for i in range(num):
        if is_outlet[i]:
            continue
        steepest = 0.0
        for (j, distance) in adj[i]:
            if elevations[i] > elevations[j]:
                slope = (elevations[i] - elevations[j]) / distance
                if slope > steepest:
                    steepest = slope
                    nxt[i] = j
    subroot = [i if is_outlet[i] else None for i in range(num)]
    has_lake = False
    for i in range(num):
        if subroot[i] is not None:
            continue
        iv = i
        while subroot[iv] is None and iv != nxt[iv]:
            iv = nxt[iv]
        if subroot[iv] is None:
            has_lake = True
            ir = iv
        else:
            ir = subroot[iv]
        iv = i
        while subroot[iv] is None and iv != nxt[iv]:
            subroot[iv] = ir
            iv = nxt[iv]
        subroot[iv] = ir
    initial = list(nxt)
    order = [None] * num
    if has_lake:
        root = [None] * num
        heap = RidgeHeap()
        for o in outlets:
            root[o] = o
            heap.push((o, 0.0))
        visited = [False] * num
        seq = 0
        while True:
            el = heap.pop()
            if el is None:
                break
            i = el[0]
            if visited[i]:
                continue
            order[i] = seq
            seq += 1
            for (j, distance) in adj[i]:
                if visited[j]:
                    continue
                if root[subroot[j]] is None:
                    k, nk = j, i
                    while nxt[k] != k:
                        tmp = nxt[k]
                        nxt[k] = nk
                        nk = k
                        k = tmp
                    nxt[k] = nk
                    root[subroot[j]] = root[subroot[i]]
                heap.push((j, distance))
            root[i] = root[subroot[i]]
            visited[i] = True
    return dict(next=nxt, next_initial=initial, subroot=subroot, has_lake=has_lake, flood_order=order)


# ---- generator.rs loop body -------------------------------------------------------------------------
def _fmax(a, b):
    """Rust f64::max: NaN loses."""
    if a != a:
        return b
    if b != b:
        return a
    return a if a > b else b


def iterate_once(adj, areas, erodibility, uplift, max_slope, outlets, elevations):
    num = len(adj)
    st = stream_tree(adj, elevations, outlets)
    nxt = st["next"]
    drainage = list(areas)
    response = [0.0] * num
    elevations = list(elevations)
    changed = False
    for outlet in outlets:
        traversal = [outlet]
        i = 0
        while True:
            it = traversal[i]
            for (jt, _) in adj[it]:
                if nxt[jt] == it:
                    traversal.append(jt)
            i += 1
            if i >= len(traversal):
                break
        for i in reversed(traversal):
            j = nxt[i]
            if j != i:
                drainage[j] += drainage[i]
        for i in traversal:
            j = nxt[i]
            ok, edge = has_edge(adj, i, j)
            distance = edge if ok else 1.0
            celerity = erodibility[i] * math.sqrt(drainage[i])  # powf(0.5), see DESIGN.md "FP discipline"
            response[i] += response[j] + 1.0 / celerity * distance
        for i in traversal:
            new_elevation = elevations[outlet] + uplift[i] * _fmax(response[i] - response[outlet], 0.0)
            ms = None if max_slope is None else max_slope[i]
            if ms is not None and not math.isnan(ms):
                j = nxt[i]
                ok, edge = has_edge(adj, i, j)
                distance = edge if ok else 1.0
                t = math.tan(ms)
                slope = (new_elevation - elevations[j]) / distance
                if slope > t:
                    new_elevation = elevations[j] + t * distance
            changed = changed or (new_elevation != elevations[i])
            elevations[i] = new_elevation
    st.update(elevations=elevations, drainage=drainage, response=response, changed=changed)
    return st


def generate(adj, areas, erodibility, uplift, max_slope, outlets, initial, max_iteration=None):
    e = list(initial)
    it = 0
    limit = 0xFFFFFFFF if max_iteration is None else max_iteration
    while it < limit:
        r = iterate_once(adj, areas, erodibility, uplift, max_slope, outlets, e)
        e = r["elevations"]
        it += 1
        if not r["changed"]:
            break
    return e, it


# ---- Terrain2D::get_elevation: Sibson's interpolant straight from its definition ----------------------
# terrain.rs:36-38 -> interpolator.rs:17-27 -> naturalneighbor::Interpolator::interpolate (crate not vendored;
# PARITY UNPINNED).  No triangulation, no cavity: Voronoi cells are built by clipping a large box with
# perpendicular-bisector half-planes, so this shares nothing with nn_oracle.cpp or the device code.
#   V(i)  = { x : |x - s_i| <= |x - s_j| for all j }          (sites only)
#   V'(p) = { x : |x - p|   <= |x - s_j| for all j }
#   w_i   = area(V'(p) & V(i)) / area(V'(p)),   z(p) = sum_i w_i z_i
# Valid for queries whose new cell is bounded by the bisectors (away from the convex hull); O(n^2) per query.
def _clip_halfplane(poly, a, b, c):
    """Keep the part of the convex polygon `poly` with a*x + b*y <= c (Sutherland-Hodgman, one plane)."""
    out = []
    k = len(poly)
    for i in range(k):
        p, q = poly[i], poly[(i + 1) % k]
        fp, fq = a * p[0] + b * p[1] - c, a * q[0] + b * q[1] - c
        if fp <= 0.0:
            out.append(p)
        if (fp < 0.0 < fq) or (fq < 0.0 < fp):
            t = fp / (fp - fq)
            out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
    return out


def _bisector(p, s):
    """Half-plane of the points at least as close to p as to s:  2 (s - p) . x <= |s|^2 - |p|^2."""
    return 2.0 * (s[0] - p[0]), 2.0 * (s[1] - p[1]), (s[0] * s[0] + s[1] * s[1]) - (p[0] * p[0] + p[1] * p[1])


def _poly_area(poly):
    if len(poly) < 3:
        return 0.0
    o = poly[0]
    acc = 0.0
    for i in range(1, len(poly) - 1):
        acc += (poly[i][0] - o[0]) * (poly[i + 1][1] - o[1]) - (poly[i][1] - o[1]) * (poly[i + 1][0] - o[0])
    return 0.5 * acc


def nn_interpolate(sites, values, p, box=1.0e4):
    """Sibson interpolation of `values` at p from the definition.  Returns (z, weights dict) ."""
    sites = [(float(x), float(y)) for x, y in sites]
    p = (float(p[0]), float(p[1]))
    cell = [(-box, -box), (box, -box), (box, box), (-box, box)]
    for s in sites:
        cell = _clip_halfplane(cell, *_bisector(p, s))
    total = _poly_area(cell)
    weights = {}
    for i, s in enumerate(sites):
        part = cell
        for j, t in enumerate(sites):
            if j != i:
                part = _clip_halfplane(part, *_bisector(s, t))
                if len(part) < 3:
                    break
        a = _poly_area(part)
        if a > 0.0:
            weights[i] = a / total
    z = math.fsum(w * float(values[i]) for i, w in weights.items())
    return z, weights
