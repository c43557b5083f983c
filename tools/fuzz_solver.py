"""Differential fuzz of the solver against the oracle on random graphs (CPU: host emulation build; with --gpu the
product library on cuda:0).
    python tools/fuzz_solver.py [--gpu] [--options] [first_seed] [count] [min_n] [max_n]
--options: instead of looping over the sweep implementations, every case runs the default sweep with a random combination
of the solver options (incremental, incr_div, rebuild_*, park_after, key_base, fuse_levels, first_flow, flood_device),
twice on the same context.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def random_graph(rng, min_n=2, max_n=80):
    """Simple symmetric graph in the boundary format: k nearest neighbours of random points plus random chords."""
    n = int(rng.integers(min_n, max_n))
    pts = rng.random((n, 2)) * 10
    edges = set()
    k = int(rng.integers(1, 5))
    for i in range(n):
        d = ((pts - pts[i]) ** 2).sum(1)
        d[i] = 1e9
        for j in np.argsort(d)[:k]:
            a, b = min(i, int(j)), max(i, int(j))
            if a != b:
                edges.add((a, b))
    for _ in range(int(rng.integers(0, n))):
        a, b = int(rng.integers(0, n)), int(rng.integers(0, n))
        if a != b:
            edges.add((min(a, b), max(a, b)))
    if n > 40 and rng.random() < 0.15:  # a hub with more than 32 neighbours (rows that do not fit a 32-bit child mask)
        hub = int(rng.integers(0, n))
        for j in rng.choice(n, int(rng.integers(33, n)), replace=False):
            if int(j) != hub:
                edges.add((min(hub, int(j)), max(hub, int(j))))
    edges = sorted(edges)
    rng.shuffle(edges)
    tie = rng.random() < 0.3  # lengths rounded to one decimal: many exact ties
    adj = [[] for _ in range(n)]
    for a, b in edges:
        w = float(np.round(np.sqrt(((pts[a] - pts[b]) ** 2).sum()), 1 if tie else 12)) or 0.1
        adj[a].append((b, w))
        adj[b].append((a, w))
    rp = np.zeros(n + 1, np.uint32)
    col, dist = [], []
    for i in range(n):
        for j, w in adj[i]:
            col.append(j)
            dist.append(w)
        rp[i + 1] = len(col)
    return dict(n=n, row_ptr=rp, col=np.array(col, np.uint32), dist=np.array(dist, np.float64), areas=rng.random(n) + 0.1)


def random_case(seed, O, min_n=2, max_n=80, max_iter=(1, 40)):
    rng = np.random.default_rng(seed)
    m = random_graph(rng, min_n, max_n)
    n = m["n"]
    p = dict(base=np.zeros(n) if rng.random() < 0.7 else rng.random(n) * 1e-3, erodibility=0.2 + rng.random(n) * 2,
             uplift=np.ones(n) if rng.random() < 0.5 else 0.5 + rng.random(n), max_slope=None)
    if rng.random() < 0.3:
        ms = 0.05 + rng.random(n) * 0.8
        ms[rng.random(n) < 0.4] = np.nan
        p["max_slope"] = ms
    outlets = np.sort(rng.choice(n, int(rng.integers(1, max(2, n // 5))), replace=False)).astype(np.uint32)
    return m, p, outlets, O.initial_elevations(p["base"]), int(rng.integers(*max_iter))


def main():
    import helpers
    from fastlem_b200 import _native, build
    from oracle import oracle as O
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    with_options = "--options" in sys.argv
    lib = _native.LIB_PATH if "--gpu" in sys.argv else build.build_emu()
    first, count = (int(args[0]) if args else 0), (int(args[1]) if len(args) > 1 else 400)
    min_n, max_n = (int(args[2]) if len(args) > 2 else 2), (int(args[3]) if len(args) > 3 else 80)
    bad = 0
    for seed in range(first, first + count):
        case = random_case(seed, O, min_n, max_n, (1, 40) if max_n <= 200 else (20, 200))
        if case is None:
            continue
        m, p, outlets, initial, mi = case
        ref, ref_it = O.generate(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, initial, mi)
        if with_options:
            rng = np.random.default_rng(seed + 77)
            opts = dict(incremental=int(rng.integers(0, 2)), incr_div=int(rng.choice([1, 2, 16, 64])),
                        rebuild_every=int(rng.choice([0, 0, 1, 2, 5, 1000])), park_after=int(rng.choice([0, 4, 8, 64])),
                        key_base=int(rng.choice([1, 2, 3, 254])), fuse_levels=int(rng.integers(0, 2)),
                        first_flow=int(rng.integers(0, 2)), flood_device=int(rng.integers(0, 2)), k5_split=int(rng.integers(0, 4) != 0), k5_cut=int(rng.integers(-1, 9)), k5_top_cap=int(rng.choice([0, 10, 200, 12288])),
                        rebuild_growth=int(rng.choice([1, 4, 50])), rebuild_height=int(rng.choice([100, 150, 400])))
            try:
                with _native.Context(0, lib) as ctx:
                    for k, v in opts.items():
                        ctx.set_option(k, v)
                    helpers.load_ctx(ctx, m, p, outlets, initial)
                    runs = [ctx.generate(mi), ctx.generate(mi)]
            except _native.FastlemError as ex:
                print("ERROR seed", seed, "n", m["n"], opts, ex)
                bad += 1
                continue
            if any(it != ref_it or not np.array_equal(e, ref, equal_nan=True) for e, it in runs):
                print("MISMATCH seed", seed, "n", m["n"], opts)
                bad += 1
            continue
        for sweep in (0, 1, 2, 3):
            with _native.Context(0, lib) as ctx:
                ctx.set_option("sweep", sweep)
                helpers.load_ctx(ctx, m, p, outlets, initial)
                e, it = ctx.generate(mi)
            if it != ref_it or not np.array_equal(e, ref, equal_nan=True):
                print("MISMATCH seed", seed, "sweep", sweep, "n", m["n"], "iterations", it, ref_it)
                bad += 1
    print("seeds", first, "..", first + count - 1, "mismatches", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
