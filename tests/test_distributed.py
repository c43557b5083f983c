"""CPU tier: the multi-GPU path (ensemble sharding + final gather) with gloo, world size 2.
Compute runs on the host emulation build here; on GPUs the same code runs with the product library and nccl."""
import os
import sys

import numpy as np
import pytest

from scenarios import ROOT, scenario


def _member(name, n, k_scale):
    m, p, outlets, initial, _ = scenario(name, n)
    return dict(row_ptr=m["row_ptr"], col=m["col"], dist=m["dist"], areas=m["areas"], initial=initial,
                erodibility=p["erodibility"] * k_scale, uplift=p["uplift"], tan_max_slope=None, outlets=outlets), m, p


def _worker(rank, world, port, emu_lib, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from fastlem_b200 import _native, ensemble
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_members = 5
    scales = [1.0, 0.5, 2.0, 1.5, 0.75]
    mine = ensemble.members_of_rank(n_members, rank, world)
    inputs = [_member("uniform", 600, scales[t])[0] for t in mine]
    res = ensemble.run_members(inputs, lambda: _native.Context(0, emu_lib), max_iteration=None)
    local = {t: res[k][0] for k, t in enumerate(mine)}
    allv = ensemble.gather_elevations(local, n_members, inputs[0]["areas"].size, rank, world)
    np.save(os.path.join(out_dir, f"gathered_{rank}.npy"), allv)
    dist.destroy_process_group()


def _raster_worker(rank, world, port, emu_lib, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from fastlem_b200 import _native, ensemble
    from tools import workloads as W
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sites, tri, he, values = _raster_inputs()
    with _native.Interpolator(sites, tri, he, lib_path=emu_lib) as it:
        it.set_values(values)
        img = ensemble.raster_partitioned(it, RASTER, rank, world)
    np.save(os.path.join(out_dir, f"raster_{rank}.npy"), img)
    dist.destroy_process_group()


def _pool_worker(rank, world, port, emu_lib, out_dir, n_contexts=1):
    """The dynamic ensemble pool (bench.py --gpus N): one shared graph per rank, members taken first come first served
    through the process group's store, results gathered at the end (variable number of members per rank)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from fastlem_b200 import _native, ensemble
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, p, outlets, initial, _ = scenario("uniform", 500)
    n = m["n"]
    n_members = 7
    pool = ensemble.MemberPool.for_process_group(n_members, "pool_test")
    mine = {}

    def make_params(t):
        return dict(initial=initial, erodibility=_pool_erodibility(n, t), uplift=p["uplift"], outlets=outlets)

    import threading
    lock = threading.Lock()

    def on_result(t, it, ctx):
        e = ctx.download()
        with lock:
            mine[t] = (e, it)
    ctxs = []
    for _ in range(n_contexts):  # members in flight per rank: one context and host thread each (bench.py --contexts-per-gpu)
        ctx = _native.Context(0, emu_lib)
        ctx.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
        ctxs.append(ctx)
    done = [t for part in ensemble.run_pool_concurrent(ctxs, pool, make_params, on_result) for t in part]
    for ctx in ctxs:
        ctx.close()
    assert sorted(done) == sorted(mine) and len(done) == len(set(done))
    # gather: member ids and elevations, padded to the largest count
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(done)], dtype=torch.int64))
    most = int(max(int(c) for c in counts))
    ids = torch.full((most,), -1, dtype=torch.int64)
    vals = torch.zeros((most, n), dtype=torch.float64)
    for k, t in enumerate(done):
        ids[k] = t
        vals[k] = torch.from_numpy(mine[t][0])
    all_ids = [torch.empty_like(ids) for _ in range(world)]
    all_vals = [torch.empty_like(vals) for _ in range(world)]
    dist.all_gather(all_ids, ids)
    dist.all_gather(all_vals, vals)
    if rank == 0:
        out = {}
        for r in range(world):
            for k in range(most):
                t = int(all_ids[r][k])
                if t >= 0:
                    assert t not in out, "a member ran twice"
                    out[t] = all_vals[r][k].numpy()
        assert sorted(out) == list(range(n_members)), "every member ran exactly once"
        np.save(os.path.join(out_dir, "pool.npy"), np.stack([out[t] for t in range(n_members)]))
    dist.destroy_process_group()


def _pool_erodibility(n, t):
    return 0.5 + np.random.default_rng(100 + t).random(n)


RASTER = dict(width=45, height=37, x0=0.0, y0=0.0, span_x=100.0, span_y=100.0, pixel_offset=0.5)


def _raster_inputs():
    from tools import workloads as W
    m = W.delaunay_model(W.random_sites(900, seed=8))
    sites, tri, he = W.triangulation_of(m)
    return sites, tri, he, 20.0 + 10.0 * W.value_noise(sites, 0.07, seed=2, octaves=3)


def test_rows_of_rank_partition():
    from fastlem_b200 import ensemble
    for height in (1, 7, 37, 4096):
        for world in (1, 2, 3, 4, 8):
            blocks = [ensemble.rows_of_rank(height, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == height
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes[0] == max(sizes)


def test_raster_two_ranks_gloo(tmp_path, emu_lib, oracle):
    """The get_elevation raster partitioned by row blocks over 2 ranks equals the single-rank image and the oracle."""
    import torch.multiprocessing as mp
    from fastlem_b200 import _native
    port = 30500 + (os.getpid() % 1000)
    mp.spawn(_raster_worker, args=(2, port, emu_lib, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "raster_0.npy")
    b = np.load(tmp_path / "raster_1.npy")
    assert np.array_equal(a, b, equal_nan=True)
    sites, tri, he, values = _raster_inputs()
    with _native.Interpolator(sites, tri, he, lib_path=emu_lib) as it:
        it.set_values(values)
        single = it.raster(it.raster_desc(RASTER["width"], RASTER["height"], 0.0, 0.0, 100.0, 100.0, 0.5))
    assert np.array_equal(a, single, equal_nan=True)
    cols, rows = np.meshgrid(np.arange(RASTER["width"]), np.arange(RASTER["height"]))
    q = np.stack([100.0 * ((cols.reshape(-1) + 0.5) / RASTER["width"]),
                  100.0 * ((rows.reshape(-1) + 0.5) / RASTER["height"])], axis=1)
    ref = oracle.nn_interpolate(sites, tri, values, q).reshape(a.shape)
    assert np.array_equal(np.isnan(a), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert (np.abs(a[ok] - ref[ok]) <= 1e-9 * np.maximum(1.0, np.abs(ref[ok]))).all()


def test_members_of_rank_partition():
    from fastlem_b200 import ensemble
    for world in (1, 2, 4, 8):
        seen = sorted(t for r in range(world) for t in ensemble.members_of_rank(64, r, world))
        assert seen == list(range(64))
        sizes = [len(ensemble.members_of_rank(64, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_ensemble_two_ranks_gloo(tmp_path, emu_lib, oracle):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, emu_lib, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "gathered_0.npy")
    b = np.load(tmp_path / "gathered_1.npy")
    assert np.array_equal(a, b), "every rank holds the same gathered ensemble"
    scales = [1.0, 0.5, 2.0, 1.5, 0.75]
    for t, sc in enumerate(scales):
        inp, m, p = _member("uniform", 600, sc)
        ref, _ = oracle.generate(m, inp["erodibility"], inp["uplift"], None, inp["outlets"], inp["initial"])
        assert np.array_equal(a[t], ref), f"member {t}"


def test_member_pool_single_process():
    from fastlem_b200 import ensemble
    pool = ensemble.MemberPool(5)
    assert [pool.take() for _ in range(7)] == [0, 1, 2, 3, 4, None, None]


@pytest.mark.parametrize("n_contexts", [1, 2])
def test_ensemble_pool_two_ranks_gloo(tmp_path, emu_lib, oracle, n_contexts):
    """Dynamic member assignment over 2 ranks (and 1 or 2 members in flight per rank): every member exactly once, each
    bit-identical to the oracle."""
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 1000) + 1000 * n_contexts
    mp.spawn(_pool_worker, args=(2, port, emu_lib, str(tmp_path), n_contexts), nprocs=2, join=True)
    got = np.load(tmp_path / "pool.npy")
    m, p, outlets, initial, _ = scenario("uniform", 500)
    for t in range(got.shape[0]):
        ref, _ = oracle.generate(m, _pool_erodibility(m["n"], t), p["uplift"], None, outlets, initial)
        assert np.array_equal(got[t], ref), f"member {t}"
