"""CPU tier: the C-ABI shared library loads and exports every symbol include/fastlem_b200.h declares;
argument validation and call-order errors (exercised through the host emulation build, no GPU compute)."""
import ctypes
import os
import re

import numpy as np
import pytest

from scenarios import ROOT, scenario

HEADER = os.path.join(ROOT, "include", "fastlem_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fastlem_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    from fastlem_b200 import _native
    assert declared_symbols() == sorted(_native.SYMBOLS)


def test_product_library_exports_every_declared_symbol(product_lib):
    lib = ctypes.CDLL(product_lib)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    lib.fastlem_version.restype = ctypes.c_char_p
    assert lib.fastlem_version().decode().endswith("sm_100a")


def test_product_library_has_sm100a_code(product_lib):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", product_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_device(product_lib):
    """On a box without a GPU the product must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from fastlem_b200 import _native
    with pytest.raises(_native.FastlemError) as ei:
        _native.Context(0, product_lib)
    assert ei.value.code == _native.E_CUDA


def test_missing_library_fails_loudly(tmp_path):
    from fastlem_b200 import _native
    with pytest.raises(ImportError):
        _native.load(str(tmp_path / "nope.so"))


def test_call_order_and_validation(emu_lib):
    from fastlem_b200 import _native
    m, p, outlets, initial, _ = scenario("uniform", 800)
    with _native.Context(0, emu_lib) as ctx:
        with pytest.raises(_native.FastlemError) as ei:  # ModelNotSet analogue
            ctx.n = m["n"]
            ctx.set_parameters(initial, p["erodibility"], p["uplift"], None, outlets)
        assert ei.value.code == _native.E_STATE
        ctx.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
        with pytest.raises(_native.FastlemError) as ei:  # ParametersNotSet analogue
            ctx.run()
        assert ei.value.code == _native.E_STATE
        with pytest.raises(ValueError):  # InvalidNumberOfParameters analogue (host mirror checks sizes)
            ctx.set_parameters(initial[:-1], p["erodibility"], p["uplift"], None, outlets)
        with pytest.raises(_native.FastlemError) as ei:
            ctx.set_parameters(initial, p["erodibility"], p["uplift"], None, np.array([m["n"]], dtype=np.uint32))
        assert ei.value.code == _native.E_INVALID
        bad_col = m["col"].copy()
        bad_col[3] = m["n"] + 7
        with pytest.raises(_native.FastlemError) as ei:
            ctx.set_graph(m["row_ptr"], bad_col, m["dist"], m["areas"])
        assert ei.value.code == _native.E_INVALID
        with pytest.raises(_native.FastlemError):
            ctx.set_option("no_such_option", 1)


def test_null_context_is_rejected(emu_lib):
    from fastlem_b200 import _native
    lib = _native.load(emu_lib)
    assert lib.fastlem_run(None, 1, None) == _native.E_INVALID
    assert lib.fastlem_create(None, 0) == _native.E_INVALID
    lib.fastlem_destroy(None)  # no-op


def test_no_outlets_runs_one_idle_iteration(emu_lib, oracle):
    """Empty outlet list: the loop body visits nothing, changed stays false (generator.rs:207-209)."""
    from fastlem_b200 import _native
    m, p, _, initial, _ = scenario("uniform", 800)
    none = np.zeros(0, dtype=np.uint32)
    with _native.Context(0, emu_lib) as ctx:
        ctx.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
        ctx.set_parameters(initial, p["erodibility"], p["uplift"], None, none)
        e, it = ctx.generate()
    ref, ref_it = oracle.generate(m, p["erodibility"], p["uplift"], None, none, initial)
    assert it == ref_it == 1 and np.array_equal(e, ref) and np.array_equal(e, initial)


def test_non_simple_graphs_are_rejected(emu_lib):
    """The solver needs what terrain-graph's add_edge produces from a triangulation: a simple, symmetric graph with
    equal lengths in both directions.  Anything else is refused by fastlem_set_graph instead of diverging later."""
    from fastlem_b200 import _native
    ok = dict(row_ptr=np.array([0, 1, 3, 4], dtype=np.uint32), col=np.array([1, 0, 2, 1], dtype=np.uint32),
              dist=np.array([1.0, 1.0, 2.0, 2.0]), areas=np.ones(3))
    cases = {
        "self loop": dict(row_ptr=np.array([0, 2, 3, 3], dtype=np.uint32), col=np.array([0, 1, 0], dtype=np.uint32),
                          dist=np.array([1.0, 1.0, 1.0])),
        "parallel edges": dict(row_ptr=np.array([0, 2, 4, 4], dtype=np.uint32), col=np.array([1, 1, 0, 0], dtype=np.uint32),
                               dist=np.array([1.0, 1.5, 1.0, 1.5])),
        "edge without its reverse": dict(row_ptr=np.array([0, 1, 1, 1], dtype=np.uint32), col=np.array([1], dtype=np.uint32),
                                         dist=np.array([1.0])),
        "lengths differ": dict(row_ptr=np.array([0, 1, 2, 2], dtype=np.uint32), col=np.array([1, 0], dtype=np.uint32),
                               dist=np.array([1.0, 2.0])),
    }
    with _native.Context(0, emu_lib) as ctx:
        ctx.set_graph(ok["row_ptr"], ok["col"], ok["dist"], ok["areas"])
        for what, g in cases.items():
            with pytest.raises(_native.FastlemError) as ei:
                ctx.set_graph(g["row_ptr"], g["col"], g["dist"], np.ones(3))
            assert ei.value.code == _native.E_INVALID, what
            assert what.split()[0] in str(ei.value), what


def test_host_graph_from_triangles_matches_builder_rule(product_lib):
    """fastlem_host_graph_from_triangles (scope row f4) = builder.rs:252-268: edge filter from < to per half-edge in
    triangle order, add_edge appends to both rows, Euclidean lengths -- against the numpy restatement in
    tools/workloads.py, bit for bit; pure host code, so it runs from the product library without a device."""
    from fastlem_b200 import _native
    from tools import workloads as W
    for n, seed in ((4, 1), (300, 2), (20000, 3)):
        m = W.delaunay_model(W.random_sites(n, seed=seed))
        rp, col, dist = _native.host_graph_from_triangles(m["sites"], m["triangles"].reshape(-1), product_lib)
        assert np.array_equal(rp, m["row_ptr"]) and np.array_equal(col, m["col"]) and np.array_equal(dist, m["dist"])
    # hand-made: two triangles (0,1,2), (2,1,3): half-edges 0->1, 1->2 | 1->3 are kept (2->0, 2->1, 3->2 are not)
    sites = np.array([[0.0, 0.0], [3.0, 0.0], [0.0, 4.0], [3.0, 4.0]])
    rp, col, dist = _native.host_graph_from_triangles(sites, np.array([0, 1, 2, 2, 1, 3], dtype=np.uint32), product_lib)
    assert rp.tolist() == [0, 1, 4, 5, 6]
    assert col.tolist() == [1, 0, 2, 3, 1, 1]
    assert dist.tolist() == [3.0, 3.0, 5.0, 4.0, 5.0, 4.0]
    with pytest.raises(_native.FastlemError):
        _native.host_graph_from_triangles(sites, np.array([0, 1, 7], dtype=np.uint32), product_lib)


def test_trim_memory_is_exported_and_harmless(emu_lib):
    """fastlem_trim_memory (include/fastlem_b200.h): returns the library's cached device memory to the driver; the host
    emulation has no pool and reports success."""
    import ctypes
    lib = ctypes.CDLL(emu_lib)
    lib.fastlem_trim_memory.argtypes = [ctypes.c_int]
    lib.fastlem_trim_memory.restype = ctypes.c_int
    assert lib.fastlem_trim_memory(0) == 0
