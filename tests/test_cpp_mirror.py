"""The C++ host mirror (fastlem_b200/host/fastlem.hpp) through examples/landscape_evolution.cpp.
CPU tier: linked against the host emulation build.  GPU tier: linked against the product library."""
import os
import subprocess

import numpy as np
import pytest

from scenarios import ROOT, scenario

SRC = os.path.join(ROOT, "examples", "landscape_evolution.cpp")


def _write_model(path, m, triangulation=None):
    with open(path, "wb") as f:
        np.array([m["n"], m["col"].size, m["default_outlets"].size], dtype=np.uint32).tofile(f)
        m["row_ptr"].astype(np.uint32).tofile(f)
        m["col"].astype(np.uint32).tofile(f)
        m["dist"].astype(np.float64).tofile(f)
        m["areas"].astype(np.float64).tofile(f)
        m["default_outlets"].astype(np.uint32).tofile(f)
        np.ascontiguousarray(m["sites"], dtype=np.float64).tofile(f)
        if triangulation is not None:
            tri, he = triangulation
            np.array([tri.size // 3], dtype=np.uint32).tofile(f)
            tri.astype(np.uint32).tofile(f)
            he.astype(np.uint32).tofile(f)


def _build(tmp_path, lib):
    exe = str(tmp_path / "landscape_evolution")
    libdir = os.path.dirname(lib)
    name = os.path.basename(lib)[3:-3]
    subprocess.check_call(["g++", "-O1", "-std=c++17", SRC, "-o", exe, f"-L{libdir}", f"-l{name}",
                           f"-Wl,-rpath,{libdir}"])
    return exe


def _run_case(tmp_path, lib, oracle, name, max_slope):
    m, p, outlets, initial, _ = scenario(name)
    exe = _build(tmp_path, lib)
    model, out = str(tmp_path / "model.bin"), str(tmp_path / "elev.bin")
    _write_model(model, m)
    args = [exe, model, out] + ([repr(max_slope)] if max_slope else [])
    res = subprocess.run(args, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    e = np.fromfile(out, dtype=np.float64)
    ms = None if not max_slope else np.full(m["n"], max_slope)
    ref, ref_it = oracle.generate(m, p["erodibility"], p["uplift"], ms, outlets, initial)
    assert f"iterations {ref_it}" in res.stdout
    assert np.array_equal(e, ref)


def _run_render(tmp_path, lib, oracle):
    """examples/landscape_evolution.rs:36-62: generate() then the get_elevation loop, through the C++ mirror."""
    from tools import workloads as W
    m = W.delaunay_model(W.random_sites(900, seed=17), lloyd=1, bound_min=(0, 0), bound_max=(100, 100))
    p = W.uniform_params(m["n"])
    sites, tri, he = W.triangulation_of(m)
    exe = _build(tmp_path, lib)
    model, out, img = str(tmp_path / "model.bin"), str(tmp_path / "elev.bin"), str(tmp_path / "image.bin")
    _write_model(model, m, (tri, he))
    size = 20
    res = subprocess.run([exe, model, out, "0", "100000", str(size), img], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    e = np.fromfile(out, dtype=np.float64)
    ref_e, _ = oracle.generate(m, p["erodibility"], p["uplift"], None, W.outlets_for(m, p),
                               oracle.initial_elevations(p["base"]))
    assert np.array_equal(e, ref_e)
    image = np.fromfile(img, dtype=np.float64).reshape(size, size)
    cols, rows = np.meshgrid(np.arange(size), np.arange(size))
    q = np.stack([100.0 * (cols.reshape(-1) / size), 100.0 * (rows.reshape(-1) / size)], axis=1)
    ref = oracle.nn_interpolate(sites, tri, e, q).reshape(size, size)
    assert np.array_equal(np.isnan(image), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert ok.sum() > 300
    assert (np.abs(image[ok] - ref[ok]) <= 1e-9 * np.maximum(1.0, np.abs(ref[ok]))).all()


def test_cpp_mirror_render_on_emulation(tmp_path, emu_lib, oracle):
    _run_render(tmp_path, emu_lib, oracle)


@pytest.mark.gpu
def test_cpp_mirror_render_on_gpu(tmp_path, product_lib, oracle):
    _run_render(tmp_path, product_lib, oracle)


def test_cpp_mirror_on_emulation(tmp_path, emu_lib, oracle):
    _run_case(tmp_path, emu_lib, oracle, "uniform", None)
    _run_case(tmp_path, emu_lib, oracle, "max_slope", 3.14 * 0.1)


@pytest.mark.gpu
def test_cpp_mirror_on_gpu(tmp_path, product_lib, oracle):
    _run_case(tmp_path, product_lib, oracle, "uniform", None)
    _run_case(tmp_path, product_lib, oracle, "max_slope", 3.14 * 0.1)
