#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: sites/s per generate() iteration.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--sites S] [--impl reference]

N = 1.  A *step* is one complete `generate()` to convergence (reference src/lem/generator.rs:140-210) on the BASELINE
config C2 workload: 1M random sites in [0,100]^2, Delaunay adjacency in the boundary format, uniform erodibility 1.0,
uplift 1.0, hull ("border") outlets.  metric = sites x iterations / seconds.
  value : device-resident (graph + parameters already in HBM, fastlem_run only; CUDA events)
  e2e   : the same through the C ABI with HOST buffers every step: fastlem_set_graph + set_parameters
          (H2D copies, flood-order prep) + fastlem_generate (D2H of the elevations)
Extra keys on the N = 1 line (not part of the metric):
  c4_16M   : BASELINE config C4, a 16M-site generate() to convergence (jittered-lattice stand-in for the relaxed Delaunay
             graph, which takes minutes to build on the host): 1 warm + 2 timed runs, stage times, roofline fractions,
             the CPU port on the first iterations of the same workload
  window   : the first --ref-iters iterations of the C2 generate() (fresh context, host buffers) -- the same window the
             CPU arm (--impl reference) times, for a like-for-like ratio
  ensemble : members of a parameter ensemble on the C2 graph through 1 and 2 contexts of one GPU
  raster   : the `Terrain2D::get_elevation` loop of the examples as a --raster^2 image (default 4096)
N > 1 (torchrun).  BASELINE config C5 style: an ensemble of independent terrains on one shared graph (every rank
builds the same C2 model and uploads it once), members differ in their noise-driven erodibility field (seed = member
index), --members-per-rank members per rank and step.  Members are handed out first come first served through the
process group's store (fastlem_b200/ensemble.py), no collective on the data path, results stay on the device, ONE NCCL
gather at the end of the timed region; "scaling": "weak".  The raster leg partitions the image rows over the ranks
and gathers the blocks on rank 0.
--impl reference: the CPU oracle (single-threaded restatement of the reference; the crate itself is Rust and cannot be
built in this image) on the C2 workload, each step = the first --ref-iters iterations (iteration 1 timed apart).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sites_per_sec_per_generate_iteration"
UNIT = "sites/s"
# algorithmic (compulsory) bytes per site per iteration, SURVEY.md 8(d) / DESIGN.md section 4
STAGE_BYTES = {"receivers": 88.125, "labels": 8.0, "area": 20.0, "elevation": 64.0}
ITER_BYTES = 180.0
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of THIS build, written by
# tools/ncu_traffic.py (absent kernel / size -> traffic null)
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")
KERNEL_STAGE = {"k_receivers_bulk": "receivers", "k_area_flow": "area", "k_incr_start": "area", "k_area_flow_long": "area",
                "k_elev_plan": "elevation", "k_elev_top": "elevation", "k_elev_low": "elevation"}


def delaunay_workload(n_sites, seed):
    from tools import workloads as W
    t0 = time.time()
    m = W.delaunay_model(W.random_sites(n_sites, (0.0, 0.0), (100.0, 100.0), seed=seed))
    p = W.uniform_params(m["n"])
    return m, p, W.outlets_for(m, p), time.time() - t0


def lattice_workload(n_sites, seed):
    from tools import workloads as W
    t0 = time.time()
    side = max(2, int(round(n_sites ** 0.5)))
    m = W.lattice_model(side, side, jitter=0.35, seed=seed)
    p = W.uniform_params(m["n"])
    return m, p, W.outlets_for(m, p), time.time() - t0


def build_workload(kind, n_sites, seed):
    return lattice_workload(n_sites, seed) if kind == "lattice" else delaunay_workload(n_sites, seed)


def workload_name(kind, n):
    if kind == "lattice":
        return (f"C4 stand-in: jittered lattice of {n} sites in [0,100]^2 (each cell split along a random diagonal), uniform "
                f"erodibility 1.0, rim outlets; step = generate() to convergence")
    return (f"C2: {n} random sites in [0,100]^2, Delaunay graph (boundary format), uniform erodibility 1.0, hull outlets; "
            f"step = generate() to convergence")


def erodibility_basis(sites):
    """Four independent noise fields over the sites (terrain_generation_advanced.rs:136-160 style fbm noise), built once per
    process; every ensemble member mixes two of them with its own angle (member_erodibility)."""
    from tools import workloads as W
    return [W.value_noise(sites, 8.0 / 75.0, seed=1000 + k, octaves=3) for k in range(4)]


def member_erodibility(basis, t):
    """Erodibility field of ensemble member t: |noise| * 4 + 0.1 with noise = a rotation by the member's own angle in the
    plane of two basis fields -- a different field for every t at the cost of one pass over the sites (the parameter
    synthesis is the caller's business, not the solver's: it must not be what the pool waits for)."""
    t = int(t)
    a = 0.61803398875 * (t + 1) * np.pi
    f = basis[t % 4] * np.cos(a) + basis[(t // 4 + 1 + t % 4) % 4] * np.sin(a)
    return np.abs(f) * 4.0 + 0.1


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, n):
    try:
        return json.load(open(TRAFFIC_FILE)).get(f"{kernel}@{n}")
    except Exception:
        return None


class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
                for k, nm in enumerate(names):
                    if r[2 + k].strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------------------------
# CPU arm
# ------------------------------------------------------------------------------------------------------------------
def ref_window_iters(args, n):
    """Iterations per step of the CPU arm (and of the GPU arm's `window` leg, which times the same window): `--ref-iters`,
    cut down so that steps + warmup of them stay near 100 s of single-thread CPU work (about 0.19 s per iteration and
    million sites, 1 s per million sites for iteration 1 with its heap flood), never below 2."""
    per_step = 100.0 / max(args.steps + args.warmup, 1)
    fit = int((per_step - 1.0e-6 * n) / (0.19e-6 * n))
    return max(2, min(args.ref_iters, fit))


def run_reference(args, rank):
    """CPU arm: the oracle port, one thread; each step = the first `ref_iters` iterations of the C2 workload.
    Iteration 1 (whose lake removal runs the heap flood) is also timed on its own, so that the steady iterations can be
    read off the line."""
    if rank != 0:
        return
    from oracle import oracle as O
    if args.gpus > 1:
        # the N > 1 arm runs the ensemble (multi_gpu): time one of its members (member 0) on the same model, the window
        # scaled down with the model size so that a step stays ~10-20 s of CPU work
        sites_arg = args.sites if args.sites != 1000000 else args.ensemble_sites
        m, p, outlets, _ = build_workload(args.workload, sites_arg, seed=1)
        p["erodibility"] = member_erodibility(erodibility_basis(m["sites"]), 0)
    else:
        m, p, outlets, _ = build_workload(args.workload, args.sites, seed=1)
    initial = O.initial_elevations(p["base"])
    n = m["n"]
    iters = ref_window_iters(args, n)

    def step(k):
        t0 = time.perf_counter()
        _, it = O.generate(m, p["erodibility"], p["uplift"], None, outlets, initial, k)
        return time.perf_counter() - t0, it
    t_first, _ = step(1)
    for _ in range(args.warmup):
        step(iters)
    tot, tot_it = 0.0, 0
    for _ in range(args.steps):
        dt, it = step(iters)
        tot += dt
        tot_it += it
    value = n * tot_it / tot
    per_step = tot / args.steps
    steady = (per_step - t_first) / max(tot_it / args.steps - 1, 1)
    sample = (f"first {iters} iterations of generate() on the {n}-site workload per step, 1 thread (nproc={os.cpu_count()}); "
              f"iteration 1 alone {t_first:.3f} s (heap flood of lake removal), later iterations {steady:.3f} s each")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": (workload_name(args.workload, n) if args.gpus <= 1 else
                                    f"C5 member: member 0 of the ensemble the {args.gpus}-GPU arm runs ({n} sites, noise-driven "
                                    f"erodibility, {args.workload} graph); the reference solves members one after the other on "
                                    f"the host whatever the number of GPUs"),
                       "sites": n, "iterations_per_step": tot_it / args.steps, "window": f"first {iters} iterations"},
            "first_iteration_seconds": t_first, "steady_iteration_seconds": steady,
            "steady_value": n / steady if steady > 0 else None,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# roofline of one profiled run
# ------------------------------------------------------------------------------------------------------------------
STAGES = ("receivers", "labels", "lakes", "order", "area", "elevation")


def roofline_of(st, n, iters, dev_ms):
    """st: stats of runs with "profile" >= 1 summed by the caller (stage ms), st["kernels"] from a "profile" = 2 run.
    The kernel the line is about is the single kernel with the largest share of the device time; `achieved` =
    algorithmic bytes of its stage per iteration / its device time per iteration."""
    peak, peak_src = peaks()
    stage_ms = {k: st["ms_" + k] for k in STAGES}
    kern = {k: v for k, v in st.get("kernels", {}).items() if v["launches"] and k in KERNEL_STAGE}
    per_it = {k: v["ms"] / max(st["kernel_iterations"], 1) for k, v in kern.items()}
    dom = max(per_it, key=per_it.get) if per_it else "k_receivers_bulk"
    stage = KERNEL_STAGE[dom]
    alg = STAGE_BYTES[stage] * n
    ms_it = per_it.get(dom, stage_ms[stage] / max(iters, 1))
    achieved = alg / (ms_it / 1e3) / 1e9 if ms_it > 0 else 0.0
    # K1: the stage events of the timed runs bracket exactly the K1 launch (plus the two small memsets before it), over
    # ALL iterations; the "profile"=2 figure is a sample of the first iterations with a synchronisation after every kernel
    k1_ms = stage_ms["receivers"] / max(iters, 1)
    k1 = STAGE_BYTES["receivers"] * n / (k1_ms / 1e3) / 1e9 if k1_ms > 0 else 0.0
    whole = ITER_BYTES * n * iters / (dev_ms / 1e3) / 1e9 if dev_ms > 0 else 0.0
    kernels = {}
    for k, v in kern.items():
        kernels[k] = {"ms_per_iteration": per_it[k], "ms_per_launch": v["ms"] / v["launches"],
                      "launches_per_iteration": v["launches"] / max(st["kernel_iterations"], 1),
                      "stage": KERNEL_STAGE[k],
                      "share_of_iteration": per_it[k] / (dev_ms / max(iters, 1)) if dev_ms > 0 else None,
                      "traffic": ncu_traffic(k, n)}
    return {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": ncu_traffic(dom, n), "peak_source": peak_src,
            "bytes_per_launch": alg, "ms_per_launch": ms_it,
            "what": f"{dom}: the {stage} stage's algorithmic bytes per iteration ({STAGE_BYTES[stage]} B x {n} sites) over the "
                    f"kernel's device time per iteration (CUDA events around every launch, \"profile\"=2 run of the same "
                    f"workload inside this process; the timed steps themselves run with stage events only)",
            "receivers_kernel": {"kernel": "k_receivers_bulk", "achieved": k1, "frac": k1 / peak, "ms_per_launch": k1_ms,
                                 "timing": "CUDA events around the K1 stage of every iteration of the timed runs",
                                 "ms_per_launch_profile2_sample": per_it.get("k_receivers_bulk"),
                                 "bytes_per_launch": STAGE_BYTES["receivers"] * n,
                                 "traffic": ncu_traffic("k_receivers_bulk", n)},
            "whole_iteration": {"achieved": whole, "frac": whole / peak, "bytes": ITER_BYTES * n,
                                "ms": dev_ms / max(iters, 1)},
            "stage_ms_per_iteration": {k: v / max(iters, 1) for k, v in stage_ms.items()},
            "kernels": kernels}


def profiled_runs(ctx, steps, max_iter):
    """`steps` runs with stage events, then one run with per-kernel events; returns (iters, dev_ms, launches, stats)."""
    iters, dev_ms, launches = 0, 0.0, 0
    acc = {"ms_" + k: 0.0 for k in STAGES}
    acc.update({"n_" + k: 0 for k in STAGES})
    last = None
    for _ in range(steps):
        it = ctx.run(max_iter)
        st = ctx.stats()
        iters += it
        dev_ms += st["ms_run"]
        launches += st["kernel_launches"]
        for k in acc:
            acc[k] += st[k]
        last = st
    return iters, dev_ms, launches, acc, last


def kernel_profile(ctx, max_iter):
    ctx.set_option("profile", 2)
    it = ctx.run(max_iter)
    st = ctx.stats()
    ctx.set_option("profile", 1)
    return st["kernels"], it


def pinned(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()


# ------------------------------------------------------------------------------------------------------------------
# extra legs of the N = 1 line
# ------------------------------------------------------------------------------------------------------------------
def c4_leg(args, local_rank):
    """BASELINE config C4: 16M sites to convergence on one B200."""
    from fastlem_b200 import _native
    m, p, outlets, t_build = lattice_workload(args.c4_sites, seed=1)
    n = m["n"]
    initial = _native.host_initial_elevations(p["base"])
    out = {"workload": workload_name("lattice", n), "sites": n, "directed_edges": int(m["col"].size),
           "workload_build_s": t_build}
    with _native.Context(local_rank) as ctx:
        ctx.set_option("profile", 1)
        t0 = time.perf_counter()
        ctx.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
        ctx.set_parameters(initial, p["erodibility"], p["uplift"], None, outlets)
        t_up = time.perf_counter() - t0
        ctx.run(args.c4_max_iter)  # warm (also computes the flood order)
        warm = ctx.stats()
        t1 = time.perf_counter()
        iters, dev_ms, launches, acc, last = profiled_runs(ctx, 2, args.c4_max_iter)
        wall = time.perf_counter() - t1
        acc["kernels"], acc["kernel_iterations"] = kernel_profile(ctx, 200 if args.c4_max_iter is None else min(200, args.c4_max_iter))
        e = ctx.download()
    out.update({"runs": "1 warm + 2 timed generate() (device-resident inputs), then 200 iterations with per-kernel events",
                "iterations_per_run": iters / 2, "generate_seconds": wall / 2, "device_ms_per_run": dev_ms / 2,
                "value": n * iters / wall, "unit": UNIT, "upload_seconds": t_up, "flood_rank_ms": warm["ms_flood_rank"],
                "flood_rank_on_device": bool(warm["flood_on_device"]), "gpu_launches": launches,
                "incremental_area_iterations": last["incremental_iterations"],
                "layout": {"rebuilds_per_run": last["rebuilds"], "nesting_levels": last["path_levels"],
                           "segments": last["paths"]},
                "roofline": roofline_of(acc, n, iters, dev_ms),
                "elevation_range": [float(e.min()), float(e.max())]})
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        tc = time.perf_counter()
        _, itc = O.generate(m, p["erodibility"], p["uplift"], None, outlets, initial, args.c4_cpu_iters)
        tc = time.perf_counter() - tc
        out["cpu_baseline"] = {"value": n * itc / tc, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": f"first {itc} iterations of the same {n}-site generate() ({tc:.1f} s, 1 thread; "
                                         f"nproc={os.cpu_count()})"}
    return out


def window_leg(args, local_rank, hm, hp, n):
    """The first --ref-iters iterations of the C2 generate(), fresh context and host buffers every step: the window
    the CPU arm times (iteration 1 with lake removal and its flood order included)."""
    from fastlem_b200 import _native
    out = pinned(np.empty(n))
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        with _native.Context(local_rank) as c2:
            c2.set_graph(hm["row_ptr"], hm["col"], hm["dist"], hm["areas"])
            c2.set_parameters(hp["initial"], hp["erodibility"], hp["uplift"], None, hp["outlets"])
            _, it = c2.generate(ref_window_iters(args, n), out=out)
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    return {"iterations": it, "seconds": t, "value": n * it / t, "unit": UNIT,
            "what": f"first {ref_window_iters(args, n)} iterations of the C2 generate() through the C ABI with host buffers, context "
                    f"creation and destruction included (median of 3) -- the window `--impl reference` times"}


def ensemble_leg(args, local_rank, m, p, outlets, initial):
    """Members of a parameter ensemble on the C2 graph through 1 and 2 contexts (streams) of one GPU: the sweeps of one
    terrain are latency-bound, so two members in flight share the GPU almost for free."""
    from fastlem_b200 import _native, ensemble
    n = m["n"]
    members = args.ensemble_members
    basis = erodibility_basis(m["sites"])
    prm = [dict(initial=initial, erodibility=member_erodibility(basis, t), uplift=p["uplift"], outlets=outlets)
           for t in range(members)]
    out = {"members": members, "sites": n, "what": "noise-driven erodibility per member (seed = member index), shared graph, "
                                                   "hull outlets; each member = set_parameters (host buffers) + run to convergence"}
    for n_ctx in (1, 2, 3, 4, 6, 8):
        ctxs = []
        for _ in range(n_ctx):
            c = _native.Context(local_rank)
            c.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
            ctxs.append(c)
        # warm: flood order + allocations
        for c in ctxs:
            c.set_parameters(prm[0]["initial"], prm[0]["erodibility"], prm[0]["uplift"], None, prm[0]["outlets"])
            c.run(3)
        pool = ensemble.MemberPool(members)
        work = [0]
        lock = threading.Lock()

        def on_result(t, it, ctx):
            with lock:
                work[0] += it
        t0 = time.perf_counter()
        ensemble.run_pool_concurrent(ctxs, pool, lambda t: prm[t], on_result)
        dt = time.perf_counter() - t0
        for c in ctxs:
            c.close()
        out[f"contexts_{n_ctx}"] = {"seconds": dt, "iterations": work[0], "value": n * work[0] / dt, "unit": UNIT}
    out["speedup_2_contexts"] = out["contexts_2"]["value"] / out["contexts_1"]["value"]
    out["like_for_like"] = (f"contexts_{args.contexts_per_gpu} is the one-GPU figure of the workload the N > 1 lines run "
                            f"({args.contexts_per_gpu} members in flight per GPU)")
    return out


def raster_leg(args, ctx_or_elev, m, rank, world, local_rank, barrier):
    """Terrain2D::get_elevation for every pixel of a size x size image of one terrain (terrain.rs:36-38; pixel
    coordinates as in examples/landscape_evolution.rs:49-50).  Sites, triangulation and elevations are replicated,
    rows are partitioned over the ranks (fastlem_b200/ensemble.py), the row blocks are gathered on rank 0 over NCCL."""
    import torch
    import torch.distributed as dist
    from fastlem_b200 import _native, ensemble
    from tools import workloads as W
    size = args.raster
    sites, tri, he = W.triangulation_of(m)
    n = sites.shape[0]
    elev = torch.empty(n, dtype=torch.float64, device="cuda")
    it = _native.Interpolator(sites, tri, he, device=local_rank)
    if rank == 0:
        if isinstance(ctx_or_elev, torch.Tensor):
            elev.copy_(ctx_or_elev)
        else:
            ctx_or_elev.download_to_device(elev.data_ptr())
    if world > 1:
        dist.broadcast(elev, src=0)
    torch.cuda.synchronize()
    it.set_values_device(elev.data_ptr())
    r0, r1 = ensemble.rows_of_rank(size, rank, world)
    max_rows = ensemble.rows_of_rank(size, 0, world)[1]
    blk = torch.zeros((max_rows, size), dtype=torch.float64, device="cuda")
    gathered = [torch.empty_like(blk) for _ in range(world)] if (world > 1 and rank == 0) else None
    desc = it.raster_desc(size, size, 0.0, 0.0, 100.0, 100.0, 0.0, r0, r1)
    torch.cuda.synchronize()  # the interpolator runs on its own stream: torch's fills of blk must have finished

    def step():
        it.raster_device(desc, blk.data_ptr())
        if world > 1:
            dist.gather(blk, gathered, dst=0)
            torch.cuda.synchronize()
        return it.stats()["ms_query_kernel"]
    for _ in range(2):
        step()
    reps, k_ms = 5, []
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        k_ms.append(step())
    barrier()
    wall = (time.perf_counter() - t0) / reps
    # through the C ABI with a pinned HOST output buffer (this rank's rows): kernel + D2H inside the timed region
    host = torch.empty((r1 - r0, size), dtype=torch.float64).pin_memory().numpy()
    it.raster(desc, out=host)
    barrier()
    t1 = time.perf_counter()
    for _ in range(reps):
        it.raster(desc, out=host)
    barrier()
    wall_host = (time.perf_counter() - t1) / reps
    st = it.stats()
    t = torch.tensor([wall, wall_host, float(np.mean(k_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    inside = float(np.isfinite(host).mean()) if host.size else 0.0
    it.close()
    if rank != 0:
        return None
    peak, _ = peaks()
    pixels = size * size
    n_tri = tri.size // 3
    # compulsory traffic of one raster: triangulation (vertices 16, neighbours 16, circumcircle 32 bytes per triangle),
    # sites 16 + values 8 bytes per site, hint grid 4 bytes per cell, 8 bytes written per pixel
    alg = 64.0 * n_tri + 24.0 * n + 4.0 * st["grid_x"] * st["grid_y"] + 8.0 * pixels
    out = {"what": f"{size}x{size} get_elevation raster of the {n}-site terrain (natural-neighbour interpolation), "
                   f"rows partitioned over {world} GPU(s), row blocks gathered on rank 0",
           "pixels": pixels, "pixels_per_s": pixels / float(t[0]), "ms_per_raster": 1e3 * float(t[0]),
           "ms_kernel_max_over_ranks": float(t[2]), "ms_per_raster_host_output": 1e3 * float(t[1]),
           "d2h_bytes_per_raster": 8 * pixels, "ms_interpolator_setup": st["ms_setup"],
           "hint_grid": [st["grid_x"], st["grid_y"]], "fraction_inside_hull_rank0_rows": inside,
           "algorithmic_bytes": alg, "achieved_gbs": alg / world / (float(t[2]) / 1e3) / 1e9 if float(t[2]) > 0 else None,
           "kernel": "k_nn_raster", "launches_per_raster": 1}
    if out["achieved_gbs"]:
        out["frac_of_hbm_peak"] = out["achieved_gbs"] / peak
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        rows = 16  # bounded sample: sixteen image rows through the CPU oracle (walk-located variant)
        cols, rws = np.meshgrid(np.arange(size, dtype=np.float64), np.arange(size // 2, size // 2 + rows, dtype=np.float64))
        q = np.stack([100.0 * (cols.reshape(-1) / size), 100.0 * (rws.reshape(-1) / size)], axis=1)
        ev = elev.cpu().numpy()
        tq = time.perf_counter()
        ref = O.nn_interpolate(sites, tri, ev, q, walk=True)
        tq = time.perf_counter() - tq
        mesh_s = O.nn_last_mesh_seconds()
        out["cpu_baseline"] = {"value": q.shape[0] / max(tq - mesh_s, 1e-9), "unit": "pixels/s", "cores": 1, "kind": "port",
                               "sample": f"{rows} rows ({q.shape[0]} pixels) of the same raster; query time only "
                                         f"({tq - mesh_s:.2f} s; building the oracle's mesh took {mesh_s:.1f} s)"}
        got = host[size // 2 - r0: size // 2 - r0 + rows].reshape(-1) if r0 <= size // 2 and size // 2 + rows <= r1 else None
        if got is not None:
            ok = np.isfinite(ref)
            out["max_rel_err_vs_oracle_on_sample"] = float((np.abs(got[ok] - ref[ok]) / np.maximum(1.0, np.abs(ref[ok]))).max())
    return out


# ------------------------------------------------------------------------------------------------------------------
# N = 1: one terrain (C2)
# ------------------------------------------------------------------------------------------------------------------
def single_gpu(args, local_rank):
    import torch
    from fastlem_b200 import _native
    m, p, outlets, t_build = build_workload(args.workload, args.sites, seed=1)
    n = m["n"]
    initial = _native.host_initial_elevations(p["base"])
    graph_bytes = m["row_ptr"].nbytes + m["col"].nbytes + m["dist"].nbytes + m["areas"].nbytes
    param_bytes = 3 * 8 * n + outlets.nbytes

    ctx = _native.Context(local_rank)
    ctx.set_option("profile", 1)
    if args.sweep is not None:
        ctx.set_option("sweep", args.sweep)
    ctx.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
    ctx.set_parameters(initial, p["erodibility"], p["uplift"], None, outlets)
    for _ in range(args.warmup):
        ctx.run(args.max_iter)

    sampler = ClockSampler(local_rank)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    iters_total, dev_ms, launches, acc, last = profiled_runs(ctx, args.steps, args.max_iter)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    value = n * iters_total / wall
    acc["kernels"], acc["kernel_iterations"] = kernel_profile(ctx, args.max_iter)

    # e2e through the C ABI with host buffers (same steps, fresh context each time); the host buffers are pinned copies of
    # the model (torch.pin_memory), handed to the C ABI as plain pointers
    hm = {k: pinned(m[k]) for k in ("row_ptr", "col", "dist", "areas")}
    hp = {"initial": pinned(initial), "erodibility": pinned(p["erodibility"]), "uplift": pinned(p["uplift"]),
          "outlets": pinned(outlets)}
    e2e_iters = 0
    out = pinned(np.empty(n))
    e2e_parts = {"create": 0.0, "set_graph": 0.0, "set_parameters": 0.0, "generate": 0.0, "destroy": 0.0}
    for step in range(-1, args.steps):  # step -1: one untimed warm-up call (first use of the library's memory pool)
        if step == 0:
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            e2e_iters = 0
            e2e_parts = {k: 0.0 for k in e2e_parts}
        ta = time.perf_counter()
        c2 = _native.Context(local_rank)
        if args.sweep is not None:
            c2.set_option("sweep", args.sweep)
        tb = time.perf_counter()
        c2.set_graph(hm["row_ptr"], hm["col"], hm["dist"], hm["areas"])
        tc = time.perf_counter()
        c2.set_parameters(hp["initial"], hp["erodibility"], hp["uplift"], None, hp["outlets"])
        td = time.perf_counter()
        _, it = c2.generate(args.max_iter, out=out)
        te_ = time.perf_counter()
        e2e_iters += it
        e2e_stats = c2.stats()
        c2.close()
        tf = time.perf_counter()
        for k, v in zip(e2e_parts, (tb - ta, tc - tb, td - tc, te_ - td, tf - te_)):
            e2e_parts[k] += v / args.steps
    torch.cuda.synchronize()
    e2e_t = time.perf_counter() - t1
    e2e_value = n * e2e_iters / e2e_t

    def extra(fn, *a):
        try:
            return fn(*a)
        except Exception as ex:  # the extras never cost the headline line
            return {"error": f"{type(ex).__name__}: {ex}"}

    window = extra(window_leg, args, local_rank, hm, hp, n)
    raster = None
    if args.raster > 0 and args.workload == "delaunay":
        raster = extra(raster_leg, args, ctx, m, 0, 1, local_rank, torch.cuda.synchronize)
    ens = extra(ensemble_leg, args, local_rank, m, p, outlets, initial) if args.ensemble_members > 0 and args.workload == "delaunay" else None

    cpu = None
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        tc = time.perf_counter()
        _, itc = O.generate(m, p["erodibility"], p["uplift"], None, outlets, initial, args.cpu_baseline_iters)
        tc = time.perf_counter() - tc
        cpu = {"value": n * itc / tc, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"first {itc} iterations of the same {n}-site generate() ({tc:.1f} s, 1 thread; nproc={os.cpu_count()})"}
    ctx.close()
    c4 = extra(c4_leg, args, local_rank) if args.c4_sites > 0 else None

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, n), "sites": n, "directed_edges": int(m["col"].size),
                       "iterations_per_step": iters_total / args.steps,
                       "l2": "working set ~170 MB per iteration > 126 MB L2; hundreds of iterations per step, no flush",
                       "max_iteration": args.max_iter, "parallelism": "single GPU", "sweep": args.sweep},
            "generate_seconds": wall / args.steps, "device_ms_per_step": dev_ms / args.steps,
            "incremental_area_iterations": last["incremental_iterations"],
            "layout": {"rebuilds_per_step": last["rebuilds"], "nesting_levels": last["path_levels"], "segments": last["paths"]},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(graph_bytes + param_bytes),
                    "d2h_bytes_per_step": int(8 * n), "seconds_per_step": e2e_t / args.steps,
                    "flood_rank_ms": e2e_stats["ms_flood_rank"], "flood_rank_on_device": bool(e2e_stats["flood_on_device"]),
                    "upload_ms": e2e_stats["ms_upload"], "seconds_by_call": e2e_parts,
                    "device_ms_in_generate": e2e_stats["ms_run"],
                    "host_buffers": "pinned host arrays handed to the C ABI as plain pointers; a fresh context per step (create, set_graph, "
                                    "set_parameters, generate, destroy), one untimed warm-up step before the timed ones"},
            "gpu_launches": int(launches), "roofline": roofline_of(acc, n, iters_total, dev_ms), "cpu_baseline": cpu,
            "clocks": clocks, "workload_build_s": t_build, "window": window, "c4_16M": c4, "ensemble": ens,
            "raster": raster}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# N > 1: an ensemble on one shared graph, members handed out dynamically
# ------------------------------------------------------------------------------------------------------------------
def multi_gpu(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from fastlem_b200 import _native, ensemble
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # the ensemble's model (`--ensemble-sites`; `--sites` if given explicitly); the same model on every rank
    sites_arg = args.sites if args.sites != 1000000 else args.ensemble_sites
    m, p, outlets, t_build = build_workload(args.workload, sites_arg, seed=1)
    n = m["n"]
    initial = _native.host_initial_elevations(p["base"])
    per_step = args.members_per_rank * world
    graph_bytes = m["row_ptr"].nbytes + m["col"].nbytes + m["dist"].nbytes + m["areas"].nbytes
    hp_shared = {"initial": pinned(initial), "uplift": pinned(p["uplift"]), "outlets": pinned(outlets)}
    basis = erodibility_basis(m["sites"])

    def make_params(t):  # host side, on the helper thread of run_pool: overlaps the previous member's solve
        return dict(initial=hp_shared["initial"], erodibility=member_erodibility(basis, t),
                    uplift=hp_shared["uplift"], outlets=hp_shared["outlets"])

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    cap = args.members_per_rank * args.steps * 2 + 4  # device buffer for this rank's results
    results = torch.zeros((cap, n), dtype=torch.float64, device="cuda")
    counters = {"iters": 0, "launches": 0, "dev_ms": 0.0, "members": []}

    lock = threading.Lock()

    def on_device(t, it, ctx):  # (called from the context's own thread)
        with lock:
            k = len(counters["members"])
            counters["members"].append(t)
        if k < cap:
            ctx.download_to_device(results[k].data_ptr())
        st = ctx.stats()
        with lock:
            counters["iters"] += it
            counters["launches"] += st["kernel_launches"]
            counters["dev_ms"] += st["ms_run"]

    # `contexts_per_gpu` members in flight per GPU: one context (own stream, own copy of the graph) and host thread each
    ctxs = []
    for _ in range(max(1, args.contexts_per_gpu)):
        c = _native.Context(local_rank)
        c.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
        ctxs.append(c)
    # warm-up: W steps of the same shape (members from their own pool), untimed
    warm_pool = ensemble.MemberPool.for_process_group(args.warmup * per_step, "fastlem_warm")
    ensemble.run_pool_concurrent(ctxs, warm_pool, make_params, lambda t, it, c: None, args.max_iter)
    counters.update({"iters": 0, "launches": 0, "dev_ms": 0.0, "members": []})

    # the collectives of the timed region once, untimed: NCCL sets up its send / receive channels on first use (1.4 s here)
    warm = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(warm, torch.zeros(1, dtype=torch.int64, device="cuda"))
    warm_out = [torch.empty((1, n), dtype=torch.float64, device="cuda") for _ in range(world)] if rank == 0 else None
    dist.gather(results[:1].contiguous(), warm_out, dst=0)
    del warm_out
    # timed region: K steps' worth of members in ONE pool (no barrier between steps), one gather at the end
    total = args.steps * per_step
    pool = ensemble.MemberPool.for_process_group(total, "fastlem_timed")
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    t0 = time.perf_counter()
    ensemble.run_pool_concurrent(ctxs, pool, make_params, on_device, args.max_iter)
    torch.cuda.synchronize()
    t_own = time.perf_counter() - t0
    mine = len(counters["members"])
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine], dtype=torch.int64, device="cuda"))
    most = int(max(int(s) for s in sizes))
    gathered = [torch.empty((most, n), dtype=torch.float64, device="cuda") for _ in range(world)] if rank == 0 else None
    dist.gather(results[:most].contiguous(), gathered, dst=0)  # the final NCCL gather of the ensemble's elevations
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([wall, t_own, counters["dev_ms"] / 1e3], dtype=torch.float64, device="cuda")
    w = torch.tensor([float(n) * counters["iters"], float(counters["launches"]), float(mine)], dtype=torch.float64, device="cuda")
    tmin = torch.tensor([t_own], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    value = float(w[0]) / float(t[0])
    for c in ctxs:
        c.close()

    # e2e: the same ensemble through the C ABI with host buffers -- the graph uploaded inside the timed region (once per
    # rank), every member = set_parameters from pinned host arrays + generate with the elevations copied to the host
    total_e2e = min(total, 2 * per_step)  # (a shorter run of the same thing: at most two steps' worth of members)
    pool2 = ensemble.MemberPool.for_process_group(total_e2e, "fastlem_e2e")
    n_ctx = max(1, args.contexts_per_gpu)
    host_out = {}
    hm = {k: pinned(m[k]) for k in ("row_ptr", "col", "dist", "areas")}
    e2e = {"iters": 0, "members": 0}

    def on_host(t_, it, c):
        c.download(out=host_out[id(c)])
        with lock:
            e2e["iters"] += it
            e2e["members"] += 1
    barrier()
    t1 = time.perf_counter()
    c2s = []
    for _ in range(n_ctx):
        c2 = _native.Context(local_rank)
        c2.set_graph(hm["row_ptr"], hm["col"], hm["dist"], hm["areas"])
        host_out[id(c2)] = pinned(np.empty(n))
        c2s.append(c2)
    t_up = time.perf_counter() - t1
    ensemble.run_pool_concurrent(c2s, pool2, make_params, on_host, args.max_iter)
    t_pool = time.perf_counter() - t1 - t_up
    for c2 in c2s:
        c2.close()
    t_close = time.perf_counter() - t1 - t_up - t_pool
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device="cuda")
    e2e_w = torch.tensor([float(n) * e2e["iters"], float(e2e["members"])], dtype=torch.float64, device="cuda")
    dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    dist.all_reduce(e2e_w, op=dist.ReduceOp.SUM)

    raster = None
    # (the raster leg needs a second pass over the triangulation on the host: only for models up to 2M sites, where it
    # costs seconds; profiles/r2p_bench_*gpu.json hold the 1M-site raster over 2 / 4 / 8 GPUs)
    if args.raster > 0 and args.workload == "delaunay" and n <= 2000000:
        try:
            raster = raster_leg(args, results[0], m, rank, world, local_rank, barrier)
        except Exception as ex:
            raster = {"error": f"{type(ex).__name__}: {ex}"}

    if rank == 0:
        members_total = int(w[2])
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * float(t[0]) / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"C5: ensemble of {per_step} members per step ({args.members_per_rank} per GPU) x {n} sites "
                                       f"on one Delaunay graph of random sites (shared by the members, uploaded once per context), noise-driven erodibility per member, hull "
                                       f"outlets; each member = generate() to convergence; members handed out first come first "
                                       f"served over the {world} ranks, one NCCL gather of the elevations at the end",
                           "sites": n, "members_per_step": per_step, "members_timed": members_total,
                           "iterations_per_member": float(w[0]) / n / max(members_total, 1),
                           "l2": f"working set ~{170 * n // 1000000} MB per member and iteration > 126 MB L2; hundreds of iterations per member, no flush",
                           "max_iteration": args.max_iter, "parallelism": f"{world} GPUs, one process each, {n_ctx} members in flight per GPU (one context and "
                                          f"stream each)",
                           "like_for_like_n1": f"`ensemble.contexts_{n_ctx}` of the N = 1 line is this workload on one GPU (the "
                                               f"N = 1 `value` is ONE C2 terrain, which cannot use more than one stream)"},
                "balance": {"slowest_rank_s": float(t[1]), "fastest_rank_s": float(tmin[0]),
                            "gather_and_barrier_s": float(t[0]) - float(t[1])},
                "e2e": {"value": float(e2e_w[0]) / float(e2e_t[0]), "unit": UNIT,
                        "h2d_bytes_per_step": int(graph_bytes * world * n_ctx * per_step / total_e2e + per_step * (3 * 8 * n + outlets.nbytes)),
                        "d2h_bytes_per_step": int(per_step * 8 * n), "seconds_per_step": float(e2e_t[0]) * per_step / total_e2e,
                        "members": int(e2e_w[1]),
                        "seconds_rank0": {"create_and_set_graph": t_up, "members": t_pool, "destroy": t_close},
                        "host_buffers": "pinned host arrays handed to the C ABI as plain pointers; graph upload inside the timed "
                                        "region (once per context)"},
                "gpu_launches": int(w[1]), "device_seconds_max_over_ranks": float(t[2]),
                "roofline": None, "cpu_baseline": None, "clocks": clocks, "workload_build_s": t_build, "raster": raster}
        peak, peak_src = peaks()
        whole = ITER_BYTES * float(w[0]) / float(t[0]) / 1e9
        line["roofline"] = {"bound": "hbm", "kernel": "whole iteration (all ranks)", "achieved": whole, "peak": peak * world,
                            "unit": "GB/s", "frac": whole / (peak * world), "traffic": None, "peak_source": peak_src,
                            "what": "180 B x sites x iterations of all members over the wall time, against N x the measured "
                                    "copy peak; the per-kernel figures are on the N = 1 line"}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--sites", type=int, default=1000000)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="delaunay", choices=["delaunay", "lattice"],
                    help="lattice: jittered lattice of ~--sites sites (C4 stand-in; builds in seconds at 16M); the raster and "
                         "ensemble legs need the Delaunay model and are skipped")
    ap.add_argument("--ref-iters", type=int, default=50, help="iterations per step of the CPU arm (and of the `window` leg)")
    ap.add_argument("--cpu-baseline-iters", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--raster", type=int, default=4096, help="side of the get_elevation raster leg (0 = skip)")
    ap.add_argument("--c4-sites", type=int, default=16000000, help="sites of the C4 leg of the N = 1 line (0 = skip)")
    ap.add_argument("--c4-max-iter", type=int, default=None)
    ap.add_argument("--c4-cpu-iters", type=int, default=5)
    ap.add_argument("--ensemble-members", type=int, default=16, help="members of the N = 1 ensemble leg (0 = skip)")
    ap.add_argument("--members-per-rank", type=int, default=8, help="N > 1: ensemble members per rank and step")
    ap.add_argument("--ensemble-sites", type=int, default=1000000,
                    help="N > 1: sites of the ensemble's model.  BASELINE config C5 has 4M-site members: "
                         "`--ensemble-sites 4000000 --workload lattice` (profiles/r2r_bench_8gpu_c5_4M.json); the default "
                         "Delaunay model of 4M sites costs minutes of Qhull per rank")
    ap.add_argument("--contexts-per-gpu", type=int, default=4,
                    help="N > 1: ensemble members in flight per GPU (one context + host thread each)")
    ap.add_argument("--sweep", type=int, default=None, help="solver option 'sweep' (DESIGN.md)")
    ap.add_argument("--max-iter", type=int, default=None,
                    help="profiling aid: stop every generate() after this many iterations (the metric is then NOT the "
                         "benchmark's; used for ncu launch lists of the same command)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        multi_gpu(args, rank, local_rank, world)
    else:
        single_gpu(args, local_rank)


if __name__ == "__main__":
    main()
