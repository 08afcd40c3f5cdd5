"""Fixed cost of one aggregate step (create -> update -> merge -> finalize -> free) per call, on a tiny batch where the
kernels themselves take nothing: python tools/step_cost.py [world] [rows_per_rank] [workload]
One thread + one kq_ctx per GPU (like tests/test_multi_gpu.py)."""
import ctypes as C, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200")); sys.path.insert(0, ROOT)
import kqgpu, bench
world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
wname = sys.argv[3] if len(sys.argv) > 3 else "cfg3"
L = kqgpu.lib()
ctxs = [kqgpu.Context(r) for r in range(world)]
idbuf = C.create_string_buffer(kqgpu.COMM_ID_BYTES)
if world > 1:
    ctxs[0].check(L.kq_comm_unique_id(ctxs[0].h, idbuf))
out = [None] * world
bar = threading.Barrier(world)


def work(r):
    ctx = ctxs[r]
    if world > 1:
        ctx.check(L.kq_comm_init(ctx.h, idbuf.raw, r, world))
    E = kqgpu.Engine(ctx)
    wl = bench.WORKLOADS[wname](rows * world)
    batch = E.generate(wl.specs(), 42, r * rows, (r + 1) * rows); ctx.sync()
    names = ["create", "update", "merge", "finalize", "free"]
    acc = {k: [] for k in names}
    for it in range(30):
        bar.wait()
        t = [time.perf_counter()]
        agg = wl.make(E); t.append(time.perf_counter())
        agg.update(batch); t.append(time.perf_counter())
        if world > 1:
            if wl.kind == "high": agg.repartition_alltoall()
            else: agg.merge_allreduce()
        t.append(time.perf_counter())
        res = agg.finalize(); t.append(time.perf_counter())
        del res, agg; t.append(time.perf_counter())
        if it >= 10:
            for k, a, b in zip(names, t[:-1], t[1:]): acc[k].append((b - a) * 1e6)
    ctx.sync()
    out[r] = {k: sorted(v)[len(v) // 2] for k, v in acc.items()}
    if world > 1:
        ctx.check(L.kq_comm_barrier(ctx.h)); ctx.check(L.kq_comm_destroy(ctx.h))


th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
[t.start() for t in th]; [t.join() for t in th]
for r in range(world):
    print(f"rank {r} median us per call:", {k: round(v) for k, v in out[r].items()}, "total", round(sum(out[r].values())))
