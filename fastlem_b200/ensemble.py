"""Ensembles of independent terrains over several GPUs.

A single terrain's receiver forest is global, so one terrain = one GPU ("replicas only").  An ensemble
(different seeds / erodibility fields / outlet masks) shards trivially: member t runs on rank t mod world,
every rank owns one fastlem context per member it runs, there is no collective on the data path, and the
elevations are gathered once at the end (NCCL all_gather over NVLink on GPUs, gloo in the CPU tests).

The `get_elevation` raster of one terrain partitions the other way: sites, elevations and triangulation are
replicated (every rank builds the same interpolator), the image is cut into contiguous row blocks, one per rank,
and the blocks are gathered once at the end.
"""
import numpy as np


def members_of_rank(n_members, rank, world):
    """Round-robin assignment: member t -> rank t % world."""
    return list(range(rank, n_members, world))


def run_members(member_inputs, make_context, max_iteration=None):
    """Run generate() for this rank's members.

    member_inputs: list of dicts with keys row_ptr, col, dist, areas, initial, erodibility, uplift, tan_max_slope,
                   outlets (the boundary format of include/fastlem_b200.h)
    make_context : callable returning a fastlem_b200._native.Context
    returns      : list of (elevations, iterations)
    """
    out = []
    for m in member_inputs:
        with make_context() as ctx:
            ctx.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
            ctx.set_parameters(m["initial"], m["erodibility"], m["uplift"], m.get("tan_max_slope"), m["outlets"])
            out.append(ctx.generate(max_iteration))
    return out


def run_members_shared_graph(graph, member_params, make_context, max_iteration=None):
    """Ensemble members that share one model (same sites and graph) and differ in their parameters only (seeds of the
    erodibility / uplift fields, outlet masks): one context, the graph is uploaded once, every member is one
    set_parameters + generate.  The flood order of lake removal is kept as long as the outlets do not change.

    graph        : dict with row_ptr, col, dist, areas
    member_params: list of dicts with initial, erodibility, uplift, tan_max_slope (optional), outlets
    returns      : list of (elevations, iterations)
    """
    out = []
    with make_context() as ctx:
        ctx.set_graph(graph["row_ptr"], graph["col"], graph["dist"], graph["areas"])
        for m in member_params:
            ctx.set_parameters(m["initial"], m["erodibility"], m["uplift"], m.get("tan_max_slope"), m["outlets"])
            out.append(ctx.generate(max_iteration))
    return out


class MemberPool:
    """A pool of ensemble members handed out one at a time, first come first served, to whichever rank is free
    (members differ in their iteration counts by +-20 %, so a static `t mod world` split leaves ranks idle).
    The shared counter lives in the process group's store (the TCPStore torchrun sets up: one `add` round trip per
    member, nothing on the data path); with a single rank it is a local counter."""

    def __init__(self, n_members, key="fastlem_pool", store=None):
        import threading
        self.n_members = int(n_members)
        self.key = key
        self.store = store
        self._next = 0
        self._lock = threading.Lock()

    @classmethod
    def for_process_group(cls, n_members, key):
        import torch.distributed as dist
        store = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            store = dist.distributed_c10d._get_default_store()
        return cls(n_members, key, store)

    def take(self):
        """Next member index, or None when the pool is empty."""
        if self.store is not None:
            t = int(self.store.add(self.key, 1)) - 1
        else:
            with self._lock:
                t = self._next
                self._next += 1
        return t if t < self.n_members else None


def run_pool(ctx, pool, make_params, on_result, max_iteration=None):
    """Run members from `pool` on one context that already holds the shared graph (set_graph done once).

    make_params(t) -> dict(initial, erodibility, uplift, tan_max_slope (optional), outlets): host arrays of member t.
                      It runs on a helper thread while the previous member is being solved (one member ahead, so the
                      host-side preparation and the next set_parameters never wait for each other).
    on_result(t, iterations, ctx): called after member t's run, before the next set_parameters (the caller downloads
                      the elevations to the host or to a device buffer there).
    Returns the list of members this context ran.
    """
    import queue
    import threading
    ready = queue.Queue()
    go = threading.Semaphore(1)  # one member is prepared ahead of the one being solved
    failure = []

    def producer():
        try:
            while True:
                go.acquire()
                t = pool.take()
                if t is None:
                    break
                ready.put((t, make_params(t)))
        except BaseException as ex:  # surfaced on the consumer side
            failure.append(ex)
        ready.put(None)

    th = threading.Thread(target=producer, daemon=True)
    th.start()
    done = []
    while True:
        item = ready.get()
        if item is None:
            break
        t, prm = item
        ctx.set_parameters(prm["initial"], prm["erodibility"], prm["uplift"], prm.get("tan_max_slope"), prm["outlets"])
        go.release()  # the next member may be prepared while this one is solved
        it = ctx.run(max_iteration)
        on_result(t, it, ctx)
        done.append(t)
    th.join()
    if failure:
        raise failure[0]
    return done


def run_pool_concurrent(contexts, pool, make_params, on_result, max_iteration=None):
    """`run_pool` on several contexts of ONE device at once, one host thread per context, all drawing from the same pool.

    The solver's sweeps are latency-bound chains: a single terrain leaves most of the GPU idle, and
    the kernels of independent members interleave on the SMs (each context has its own stream; the library calls release
    the GIL).  `on_result(t, iterations, ctx)` is called from the context's thread -- guard shared state with a lock.
    Returns the list of members run, per context."""
    import threading
    done = [None] * len(contexts)
    failure = []

    def work(k):
        try:
            done[k] = run_pool(contexts[k], pool, make_params, on_result, max_iteration)
        except BaseException as ex:
            failure.append(ex)

    threads = [threading.Thread(target=work, args=(k,), daemon=True) for k in range(1, len(contexts))]
    for th in threads:
        th.start()
    work(0)
    for th in threads:
        th.join()
    if failure:
        raise failure[0]
    return done


def gather_elevations(local, n_members, n_sites, rank, world, device="cpu"):
    """All-gather the members' elevations: returns an (n_members, n_sites) float64 array on every rank.

    local: dict member index -> elevations (numpy) for the members this rank ran.
    Uses torch.distributed when world > 1 (the caller has initialised the process group: nccl on GPUs, gloo on CPU).
    """
    import torch
    per_rank = (n_members + world - 1) // world
    buf = torch.zeros((per_rank, n_sites), dtype=torch.float64, device=device)
    mine = members_of_rank(n_members, rank, world)
    for k, t in enumerate(mine):
        buf[k].copy_(torch.from_numpy(np.ascontiguousarray(local[t])))
    if world == 1:
        gathered = [buf]
    else:
        import torch.distributed as dist
        gathered = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(gathered, buf)
    out = np.empty((n_members, n_sites), dtype=np.float64)
    for r in range(world):
        g = gathered[r].cpu().numpy()
        for k, t in enumerate(members_of_rank(n_members, r, world)):
            out[t] = g[k]
    return out


def rows_of_rank(height, rank, world):
    """Contiguous row block [begin, end) of a `height`-row raster for `rank`: sizes differ by at most one row."""
    base, extra = divmod(height, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def raster_partitioned(interpolator, desc_args, rank, world, device="cpu"):
    """Every rank rasterises its own row block and all ranks end up with the full (height, width) image.

    interpolator: fastlem_b200._native.Interpolator with values set (replicated on every rank)
    desc_args   : dict(width, height, x0, y0, span_x, span_y, pixel_offset)
    On GPUs (`device` = "cuda:k") the block is written straight into the torch tensor that NCCL gathers
    (fastlem_interp_raster_device, no host round trip); with gloo the block goes through the host.
    """
    import torch
    width, height = int(desc_args["width"]), int(desc_args["height"])
    r0, r1 = rows_of_rank(height, rank, world)
    max_rows = rows_of_rank(height, 0, world)[1]
    buf = torch.zeros((max_rows, width), dtype=torch.float64, device=device)
    desc = interpolator.raster_desc(width, height, desc_args["x0"], desc_args["y0"], desc_args["span_x"],
                                    desc_args["span_y"], desc_args.get("pixel_offset", 0.0), r0, r1)
    if r1 > r0:
        if buf.is_cuda:
            torch.cuda.synchronize(buf.device)  # the interpolator has its own stream: wait for torch's fill of buf
            interpolator.raster_device(desc, buf.data_ptr())
        else:
            buf[:r1 - r0].copy_(torch.from_numpy(interpolator.raster(desc)))
    if world == 1:
        gathered = [buf]
    else:
        import torch.distributed as dist
        gathered = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(gathered, buf)
    out = np.empty((height, width), dtype=np.float64)
    for r in range(world):
        a, b = rows_of_rank(height, r, world)
        out[a:b] = gathered[r][:b - a].cpu().numpy()
    return out
