#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: sites/s per generate() iteration.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--sites S] [--impl reference]

A *step* is one complete `generate()` to convergence (reference src/lem/generator.rs:140-210) on the
BASELINE config C2 workload: 1M random sites in [0,100]^2, Delaunay adjacency in the boundary format,
uniform erodibility 1.0, uplift 1.0, hull ("border") outlets.  metric = sites x iterations / seconds.
  value : device-resident (graph + parameters already in HBM, fastlem_run only; CUDA events)
  e2e   : the same through the C ABI with HOST buffers every step: fastlem_set_graph + set_parameters
          (H2D copies, flood-order prep) + fastlem_generate (D2H of the elevations)
N > 1 (torchrun): one independent terrain per rank (an ensemble member with its own seed), no data-path
collective, one final NCCL all_gather of the elevations per step; "scaling": "weak".
Extra key "raster" (not part of the metric): the `Terrain2D::get_elevation` loop of the examples as a --raster^2
image (default 4096) of rank 0's terrain, rows partitioned over the ranks, one NCCL all_gather of the row blocks.
--impl reference: the CPU oracle (single-threaded restatement of the reference; the crate itself is Rust and
cannot be built in this image) on the same workload, each step bounded to the first few iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sites_per_sec_per_generate_iteration"
UNIT = "sites/s"
# algorithmic bytes per site per iteration (DESIGN.md "Roofline")
STAGE_BYTES = {"receivers": 88.125, "labels": 8.0, "area": 20.0, "elevation": 64.0}
# dram__bytes_read.sum + dram__bytes_write.sum per pass at 1M sites from one `ncu --set full` capture of a late iteration
# (profiles/r1b_ncu_kernels.txt): K1 = k_receivers_mask; K4 = the two flow kernels of an incremental pass
NCU_TRAFFIC = {"receivers": 107.6e6, "area": 11.0e6}
NCU_TRAFFIC_NOTE = {"receivers": "ncu: 99.1 MB read + 8.5 MB written per launch (88.1 MB algorithmic)",
                    "area": "ncu: k_incr_start 6.8 MB + k_area_flow_long 4.1 MB per incremental pass (20 MB algorithmic for a "
                            "full pass; the incremental pass touches 3-10 % of the sites)"}


WORKLOAD = "delaunay"  # --workload lattice: jittered lattice (stand-in for the relaxed Delaunay graph of C4 at 16M sites)


def build_workload(n_sites, seed):
    from tools import workloads as W
    t0 = time.time()
    if WORKLOAD == "lattice":
        side = max(2, int(round(n_sites ** 0.5)))
        m = W.lattice_model(side, side, jitter=0.35, seed=seed)
    else:
        m = W.delaunay_model(W.random_sites(n_sites, (0.0, 0.0), (100.0, 100.0), seed=seed))
    p = W.uniform_params(m["n"])
    outlets = W.outlets_for(m, p)
    return m, p, outlets, time.time() - t0


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def run_reference(args, rank, world):
    """CPU arm: the oracle port, one thread, each step = the first `ref_iters` iterations of the workload."""
    if rank != 0:
        return
    from oracle import oracle as O
    m, p, outlets, _ = build_workload(args.sites, seed=1)
    initial = O.initial_elevations(p["base"])
    n = m["n"]
    iters = args.ref_iters

    def step():
        t0 = time.perf_counter()
        _, it = O.generate(m, p["erodibility"], p["uplift"], None, outlets, initial, iters)
        return time.perf_counter() - t0, it
    for _ in range(args.warmup):
        step()
    tot, tot_it = 0.0, 0
    for _ in range(args.steps):
        dt, it = step()
        tot += dt
        tot_it += it
    value = n * tot_it / tot
    sample = f"first {iters} iterations of generate() on the {n}-site workload per step (incl. the heap flood of iteration 1)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": (f"C2: {n} random sites, Delaunay graph, uniform erodibility, hull outlets"
                                    if WORKLOAD == "delaunay" else f"C4 stand-in: jittered lattice of {n} sites, uniform "
                                    f"erodibility, rim outlets"),
                       "sites": n, "iterations_per_step": iters},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def raster_leg(args, ctx, m, rank, world, local_rank, barrier):
    """Terrain2D::get_elevation for every pixel of a size x size image of rank 0's terrain (terrain.rs:36-38; pixel
    coordinates as in examples/landscape_evolution.rs:49-50).  Sites, triangulation and elevations are replicated,
    rows are partitioned over the ranks (fastlem_b200/ensemble.py), the row blocks are gathered once over NCCL."""
    import torch
    import torch.distributed as dist
    from fastlem_b200 import _native, ensemble
    from tools import workloads as W
    size = args.raster
    if world > 1 and rank != 0:
        m = build_workload(args.sites, seed=1)[0]  # replicate rank 0's model (host-side graph build)
    sites, tri, he = W.triangulation_of(m)
    n = sites.shape[0]
    elev = torch.empty(n, dtype=torch.float64, device="cuda")
    it = _native.Interpolator(sites, tri, he, device=local_rank)
    if rank == 0:
        ctx.download_to_device(elev.data_ptr())
    if world > 1:
        dist.broadcast(elev, src=0)
    torch.cuda.synchronize()
    it.set_values_device(elev.data_ptr())
    r0, r1 = ensemble.rows_of_rank(size, rank, world)
    max_rows = ensemble.rows_of_rank(size, 0, world)[1]
    blk = torch.zeros((max_rows, size), dtype=torch.float64, device="cuda")
    gathered = [torch.empty_like(blk) for _ in range(world)] if world > 1 else None
    desc = it.raster_desc(size, size, 0.0, 0.0, 100.0, 100.0, 0.0, r0, r1)
    torch.cuda.synchronize()  # the interpolator runs on its own stream: torch's fills of blk must have finished

    def step():
        it.raster_device(desc, blk.data_ptr())
        if world > 1:
            dist.all_gather(gathered, blk)
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
    pixels = size * size
    n_tri = tri.size // 3
    # compulsory traffic of one raster: triangulation (vertices 16, neighbours 16, circumcircle 32 bytes per triangle),
    # sites 16 + values 8 bytes per site, hint grid 4 bytes per cell, 8 bytes written per pixel
    alg = 64.0 * n_tri + 24.0 * n + 4.0 * st["grid_x"] * st["grid_y"] + 8.0 * pixels
    out = {"what": f"{size}x{size} get_elevation raster of the {n}-site terrain (natural-neighbour interpolation), "
                   f"rows partitioned over {world} GPU(s)",
           "pixels": pixels, "pixels_per_s": pixels / float(t[0]), "ms_per_raster": 1e3 * float(t[0]),
           "ms_kernel_max_over_ranks": float(t[2]), "ms_per_raster_host_output": 1e3 * float(t[1]),
           "d2h_bytes_per_raster": 8 * pixels, "ms_interpolator_setup": st["ms_setup"],
           "hint_grid": [st["grid_x"], st["grid_y"]], "fraction_inside_hull_rank0_rows": inside,
           "algorithmic_bytes": alg, "achieved_gbs": alg / world / (float(t[2]) / 1e3) / 1e9 if float(t[2]) > 0 else None,
           "kernel": "k_nn_raster", "launches_per_raster": 1}
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--sites", type=int, default=1000000)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="delaunay", choices=["delaunay", "lattice"],
                    help="lattice: jittered lattice of ~--sites sites (C4 stand-in; builds in seconds at 16M); the raster leg "
                         "needs a Delaunay triangulation and is skipped")
    ap.add_argument("--ref-iters", type=int, default=5, help="iterations per step of the CPU arm")
    ap.add_argument("--cpu-baseline-iters", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--raster", type=int, default=4096, help="side of the get_elevation raster leg (0 = skip)")
    ap.add_argument("--sweep", type=int, default=None, help="solver option 'sweep' (DESIGN.md)")
    ap.add_argument("--max-iter", type=int, default=None,
                    help="profiling aid: stop every generate() after this many iterations (the metric is then NOT the "
                         "benchmark's; used for ncu launch lists of the same command)")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    if WORKLOAD == "lattice":
        args.raster = 0

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from fastlem_b200 import _native
    m, p, outlets, t_build = build_workload(args.sites, seed=1 + rank)
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
    gathered = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(world)] if world > 1 else None
    mine = torch.empty(n, dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        it = ctx.run(args.max_iter)
        if world > 1:
            ctx.download_to_device(mine.data_ptr())
            dist.all_gather(gathered, mine)
            torch.cuda.synchronize()
        return it, ctx.stats()

    for _ in range(args.warmup):
        resident_step()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    t0 = time.perf_counter()
    iters_total, dev_ms, launches = 0, 0.0, 0
    stage_ms = {"receivers": 0.0, "labels": 0.0, "lakes": 0.0, "order": 0.0, "area": 0.0, "elevation": 0.0}
    stage_n = dict.fromkeys(stage_ms, 0)
    last = None
    for _ in range(args.steps):
        it, st = resident_step()
        iters_total += it
        dev_ms += st["ms_run"]
        launches += st["kernel_launches"]
        for k in stage_ms:
            stage_ms[k] += st["ms_" + k]
            stage_n[k] += st["n_" + k]
        last = st
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None

    # max over ranks of the bracketed time; sum of site-iterations over ranks
    t = torch.tensor([wall, dev_ms / 1e3], dtype=torch.float64, device="cuda")
    w = torch.tensor([float(n) * iters_total, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    wall_max, dev_max = float(t[0]), float(t[1])
    work, launches_all = float(w[0]), int(w[1])
    value = work / wall_max

    # e2e through the C ABI with host buffers (same steps, fresh context each time); the host buffers are
    # pinned copies of the model (torch.pin_memory), handed to the C ABI as plain pointers
    def pinned(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    hm = {k: pinned(m[k]) for k in ("row_ptr", "col", "dist", "areas")}
    hp = {"initial": pinned(initial), "erodibility": pinned(p["erodibility"]), "uplift": pinned(p["uplift"]),
          "outlets": pinned(outlets)}
    e2e_iters, e2e_t = 0, 0.0
    out = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
    barrier()
    t1 = time.perf_counter()
    e2e_parts = {"create": 0.0, "set_graph": 0.0, "set_parameters": 0.0, "generate": 0.0, "destroy": 0.0}
    for _ in range(args.steps):
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
    barrier()
    e2e_t = time.perf_counter() - t1
    te = torch.tensor([e2e_t], dtype=torch.float64, device="cuda")
    we = torch.tensor([float(n) * e2e_iters], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(we, op=dist.ReduceOp.SUM)
    e2e_value = float(we[0]) / float(te[0])

    raster = None
    if args.raster > 0:
        try:
            raster = raster_leg(args, ctx, m, rank, world, local_rank, barrier)
        except Exception as ex:  # the raster is an extra: never lose the headline line over it
            raster = {"error": f"{type(ex).__name__}: {ex}"}

    if rank == 0:
        peak, peak_src = peaks()
        dom = max(("receivers", "area", "elevation"), key=lambda k: stage_ms[k])
        # one "launch" of a stage = one pass of that stage over all n sites (= one iteration's worth)
        passes = iters_total
        alg_bytes = STAGE_BYTES[dom] * n
        achieved = alg_bytes * passes / (stage_ms[dom] / 1e3) / 1e9 if stage_ms[dom] > 0 else 0.0
        k1 = STAGE_BYTES["receivers"] * n * passes / (stage_ms["receivers"] / 1e3) / 1e9 if stage_ms["receivers"] else 0.0
        iter_bytes = 180.0 * n
        whole = iter_bytes * passes / (dev_ms / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": NCU_TRAFFIC.get(dom) if n == 1000000 else None,
                    "peak_source": peak_src,
                    "bytes_per_pass": alg_bytes, "ms_per_pass": stage_ms[dom] / max(passes, 1),
                    "kernel_launches_per_pass": stage_n[dom] / max(passes, 1),
                    "kernel_names": {"receivers": "k_receivers_mask",
                                     "area": "incremental pass (9 of 10 iterations): k_seg_keys+scan+k_incr_mark+k_incr_prepare+"
                                             "k_incr_start+k_area_flow_long+k_incr_cleanup; full pass: k_count_waits+k_seg_keys+scan+"
                                             "k_seg_prepare+k_area_flow+k_area_flow_long",
                                     "elevation": "k_celerity_term+k_fused_index+k_elev_flow_fused+k_elev_flow"}[dom],
                    "traffic_note": NCU_TRAFFIC_NOTE.get(dom),
                    "receivers_kernel": {"achieved": k1, "frac": k1 / peak,
                                         "ms_per_launch": stage_ms["receivers"] / max(stage_n["receivers"], 1)},
                    "whole_iteration": {"achieved": whole, "frac": whole / peak, "bytes": iter_bytes},
                    "stage_ms_per_iteration": {k: v / max(passes, 1) for k, v in stage_ms.items()}}
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import oracle as O
            k_it = args.cpu_baseline_iters
            tc = time.perf_counter()
            _, itc = O.generate(m, p["erodibility"], p["uplift"], None, outlets, initial, k_it)
            tc = time.perf_counter() - tc
            cpu = {"value": n * itc / tc, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"first {itc} iterations of the same {n}-site generate() ({tc:.1f} s, 1 thread; "
                             f"nproc={os.cpu_count()})"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * wall_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": (f"C2: {n} random sites in [0,100]^2, Delaunay graph (boundary format), uniform "
                                        f"erodibility 1.0, hull outlets; step = generate() to convergence"
                                        if WORKLOAD == "delaunay" else
                                        f"C4 stand-in: jittered lattice of {n} sites in [0,100]^2 (each cell split along a random "
                                        f"diagonal), uniform erodibility 1.0, rim outlets; step = generate() to convergence"),
                           "sites": n, "directed_edges": int(m["col"].size),
                           "iterations_per_step": iters_total / args.steps,
                           "l2": "working set ~170 MB per iteration > 126 MB L2; hundreds of iterations per step, no flush",
                           "max_iteration": args.max_iter,
                           "parallelism": "1 terrain per GPU" if world > 1 else "single GPU",
                           "sweep": args.sweep},
                "generate_seconds": wall_max / args.steps, "device_ms_per_step": 1e3 * dev_max / args.steps,
                "first_iteration_nesting_levels_or_tree_depth": last["depth_first"] or None,
                "incremental_area_iterations": last["incremental_iterations"],
                "layout": {"rebuilds_per_step": last["rebuilds"], "nesting_levels": last["path_levels"],
                           "segments": last["paths"]},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(graph_bytes + param_bytes),
                        "d2h_bytes_per_step": int(8 * n), "seconds_per_step": float(te[0]) / args.steps,
                        "flood_rank_ms": e2e_stats["ms_flood_rank"], "flood_rank_on_device": bool(e2e_stats["flood_on_device"]), "upload_ms": e2e_stats["ms_upload"],
                        "seconds_by_call": e2e_parts, "device_ms_in_generate": e2e_stats["ms_run"],
                        "host_buffers": "pinned host arrays handed to the C ABI as plain pointers"},
                "gpu_launches": launches_all, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
                "workload_build_s": t_build, "raster": raster}
        if raster and "achieved_gbs" in raster:
            raster["frac_of_hbm_peak"] = raster["achieved_gbs"] / peak
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
