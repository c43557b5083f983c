"""Shared comparison code: run a solver Context on a scenario and compare every stage with the oracle."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL_TOL = 1e-9  # north_star: drainage areas and elevations within 1e-9 relative; integer stages bit-exact


def tan_of(max_slope):
    """tan(max_slope) with libm's tan (math.tan), NaN = None -- what the host mirrors pass to the C ABI."""
    import math
    if max_slope is None:
        return None
    return np.array([math.tan(v) if v == v else np.nan for v in max_slope], dtype=np.float64)


def load_ctx(ctx, m, p, outlets, initial):
    ctx.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
    ctx.set_parameters(initial, p["erodibility"], p["uplift"], tan_of(p["max_slope"]), outlets)


def assert_close(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, what
    nan = np.isnan(a) | np.isnan(b)
    assert np.array_equal(np.isnan(a), np.isnan(b)), f"{what}: NaN pattern differs"
    denom = np.maximum(np.abs(b[~nan]), 1e-300)
    err = np.abs(a[~nan] - b[~nan]) / denom
    worst = float(err.max()) if err.size else 0.0
    assert worst <= REL_TOL, f"{what}: max relative error {worst:.3e} > {REL_TOL}"
    return worst


def check_first_iteration(ctx, O, m, p, outlets, initial):
    """Iteration 1, stage by stage.  Integer stages bit-exact; f64 stages within REL_TOL (and reported exact)."""
    ref = O.iterate_once(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, initial)
    ctx.set_option("keep_stages", 1)
    _, it = ctx.generate(1)
    assert it == 1
    if ref["has_lake"]:
        assert np.array_equal(ctx.fetch("receivers_initial"), ref["next_initial"]), "receivers before lake removal"
        assert np.array_equal(ctx.fetch("labels_initial"), ref["subroot"]), "basin labels (subroot)"
        assert ctx.stats()["lake_iterations"] == 1
    else:
        assert ctx.stats()["lake_iterations"] == 0
    assert np.array_equal(ctx.fetch("receivers"), ref["next"]), "receivers after lake removal"
    reached = ref["order"] != O.NONE
    depth = ctx.fetch("depth")
    assert np.array_equal(depth != 0xFFFFFFFF, reached), "set of sites visited by the per-outlet traversal"
    assert_close(ctx.fetch("drainage_area"), ref["drainage"], "drainage area")
    assert_close(ctx.fetch("response_time"), ref["response"], "response time")
    assert_close(ctx.fetch("elevation"), ref["elevations"], "elevation after iteration 1")
    exact = all(np.array_equal(ctx.fetch(s), ref[k], equal_nan=True) for s, k in
                (("drainage_area", "drainage"), ("response_time", "response"), ("elevation", "elevations")))
    return ref, exact


def check_generate(ctx, O, m, p, outlets, initial, max_iteration):
    ref_e, ref_it = O.generate(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, initial, max_iteration)
    e, it = ctx.generate(max_iteration)
    assert it == ref_it, f"iterations to convergence: {it} vs oracle {ref_it}"
    assert_close(e, ref_e, "final elevations")
    return np.array_equal(e, ref_e, equal_nan=True)


def golden_cases():
    """generate() cases (nn_*.npz are the get_elevation cases)."""
    return [p for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))) if not os.path.basename(p).startswith("nn_")]


def nn_golden_cases():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "nn_*.npz")))


def load_golden(path):
    g = np.load(path)
    m = dict(n=g["row_ptr"].size - 1, row_ptr=g["row_ptr"], col=g["col"], dist=g["dist"], areas=g["areas"])
    p = dict(base=g["base"], erodibility=g["erodibility"], uplift=g["uplift"],
             max_slope=g["max_slope"] if bool(g["has_max_slope"]) else None)
    mi = int(g["max_iteration"])
    return g, m, p, g["outlets"], (None if mi < 0 else mi)


def check_against_golden(ctx, path):
    g, m, p, outlets, max_iteration = load_golden(path)
    load_ctx(ctx, m, p, outlets, g["initial"])
    ctx.set_option("keep_stages", 1)
    ctx.generate(1)
    assert np.array_equal(ctx.fetch("receivers"), g["it1_next"])
    if bool(g["it1_has_lake"]):
        assert np.array_equal(ctx.fetch("receivers_initial"), g["it1_next_initial"])
        assert np.array_equal(ctx.fetch("labels_initial"), g["it1_subroot"])
        rank = ctx.fetch("flood_rank")
        assert np.array_equal(rank, g["it1_flood_order"])
    assert_close(ctx.fetch("drainage_area"), g["it1_drainage"], "golden drainage")
    assert_close(ctx.fetch("response_time"), g["it1_response"], "golden response")
    assert_close(ctx.fetch("elevation"), g["it1_elevations"], "golden elevation it1")
    e, it = ctx.generate(max_iteration)
    assert it == int(g["iterations"])
    assert_close(e, g["final"], "golden final elevations")
    assert np.array_equal(e, g["final"], equal_nan=True), "golden final elevations (bit-exact)"


def check_converged_properties(ctx, O, m, p, outlets, initial, first_iterations=3):
    """Size-independent checks for workloads the oracle cannot run to convergence: the first iterations against the
    oracle, then properties of the converged device result (determinism, a forest draining to the outlets, elevations
    increasing upstream, outlets untouched, conservation of drainage area, fixed point under one oracle iteration)."""
    n = m["n"]
    load_ctx(ctx, m, p, outlets, initial)
    check_first_iteration(ctx, O, m, p, outlets, initial)
    if first_iterations:
        assert check_generate(ctx, O, m, p, outlets, initial, first_iterations)
    e, it = ctx.generate()
    recv = ctx.fetch("receivers").astype(np.int64)
    A = ctx.fetch("drainage_area")
    depth = ctx.fetch("depth")
    labels = ctx.fetch("labels").astype(np.int64)
    e2, it2 = ctx.generate()
    assert it == it2 and np.array_equal(e, e2), "deterministic"
    is_outlet = np.zeros(n, dtype=bool)
    is_outlet[outlets] = True
    # converged forest: every site drains to an outlet, elevation strictly increases upstream (positive uplift)
    assert (depth != 0xFFFFFFFF).all()
    assert is_outlet[labels].all()
    non_out = ~is_outlet
    assert (recv[non_out] != np.arange(n)[non_out]).all()
    assert (e[non_out] > e[recv[non_out]]).all()
    assert np.array_equal(e[is_outlet], initial[is_outlet]), "outlets keep base + noise (generator.rs:177-179)"
    # conservation: the outlets' drainage areas add up to the total cell area
    assert abs(A[is_outlet].sum() - m["areas"].sum()) <= 1e-9 * m["areas"].sum()
    # fixed point: one more body from the converged field changes nothing (checked with the oracle)
    nxt = O.iterate_once(m, p["erodibility"], p["uplift"], p["max_slope"], outlets, e)
    assert not nxt["changed"]
    assert np.array_equal(nxt["next"], recv)
    return e, it
