"""Critical path of one K4 pass from the per-segment time stamps of a -DFL_FLOW_STATS build (tools/_dbg).
   python tools/flow_timeline.py SITES ITER [opt=value ...]"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastlem_b200 import _native
from tools import workloads as W
LIB = os.path.join(ROOT, "tools", "_dbg", "libfastlem_stats.so")

def raw(ctx, stage, n, dtype, per=1):
    out = np.zeros(n * per, dtype=dtype)
    ctx._ck(ctx._lib.fastlem_debug_fetch(ctx._h, stage, out.ctypes.data_as(ctypes.c_void_p), out.nbytes))
    return out

n = int(sys.argv[1]); it = int(sys.argv[2])
cache = f"/tmp/fl_workload_{n}_0.npz"
if os.path.exists(cache):
    z = np.load(cache); m = {k: z[k] for k in z.files}; m["n"] = n
else:
    m = W.delaunay_model(W.random_sites(n, seed=1))
p = W.uniform_params(n)
initial = _native.host_initial_elevations(p["base"], LIB)
with _native.Context(0, LIB) as ctx:
    for k, v in [a.split("=") for a in sys.argv[3:]]:
        ctx.set_option(k, int(v))
    ctx.set_graph(m["row_ptr"], m["col"], m["dist"], m["areas"])
    ctx.set_parameters(initial, p["erodibility"], p["uplift"], None, m["default_outlets"])
    ctx.run(it - 1); raw(ctx, 100, n, np.uint64, 4)          # clears the log
    ctx.run(it)                                               # the log now holds iterations 1..it; the last writer wins
    st = ctx.stats()
    t = raw(ctx, 100, n, np.uint64, 4).reshape(n, 4).astype(np.int64)
    recv = raw(ctx, 101, n, np.uint32).astype(np.int64)
    sh = raw(ctx, 102, n, np.uint32).astype(np.int64)
print("incremental iterations", st["incremental_iterations"], "of", st["iterations"])
# only segments touched in the LAST iteration: end stamp within the last pass = after the latest (min start) cluster
end = t[:, 3]
live = end > 0
t_last_end = end.max()
# the last pass is at most a few ms long
recent = live & (end > t_last_end - 1_000_000)
heads = np.nonzero(recent)[0]
t0 = min(t[heads, 0][t[heads, 0] > t_last_end - 1_000_000].min(), end[heads].min())
print("segments finished in the last pass:", heads.size, " span us:", (t_last_end - t0) / 1e3)
# tail positions per head: segment length
order = np.argsort(sh, kind="stable")
seglen = np.bincount(sh, minlength=n)
def rel(x): return (x - t0) / 1e3
cur = heads[np.argmax(end[heads])]
print(f"{'head':>9s} {'len':>5s} {'start':>9s} {'park':>9s} {'resume':>9s} {'end':>9s} {'climb_us':>9s} {'gap_from_child':>10s}")
chain = []
while True:
    s0, pk, rs, e = t[cur]
    # children segments: heads h with seg_head[recv[h]] == cur, finished in this pass
    kids = heads[(sh[recv[heads]] == cur) & (heads != cur) & (recv[heads] != heads)]
    kids = kids[end[kids] <= max(s0, 1 << 62)] if kids.size else kids
    last_kid = kids[np.argmax(end[kids])] if kids.size else -1
    gap = rel(s0) - rel(end[last_kid]) if last_kid >= 0 and s0 > 0 else float("nan")
    print(f"{cur:9d} {seglen[cur]:5d} {rel(s0) if s0>0 else -1:9.1f} {rel(pk) if pk>t0 else -1:9.1f} {rel(rs) if rs>t0 else -1:9.1f} {rel(e):9.1f} {(e-s0)/1e3 if s0>0 else -1:9.1f} {gap:10.1f}  kids={kids.size}")
    if last_kid < 0 or len(chain) > 60: break
    chain.append(cur); cur = last_kid
