import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200")); sys.path.insert(0, ROOT)
import kqgpu, bench
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
ctx = kqgpu.Context(0); E = kqgpu.Engine(ctx)
wl = bench.WORKLOADS["cfg2f"](rows)
batch = E.generate(wl.specs(), 42, 0, rows); ctx.sync()
for _ in range(3):
    r = wl.run(E, batch); del r
ctx.sync()
for mode in ("keep", "del"):
    ts = []
    ctx.timer_begin(); t0 = time.perf_counter(); keep = None
    for i in range(8):
        a = time.perf_counter()
        if mode == "keep":
            keep = wl.run(E, batch)
        else:
            r = wl.run(E, batch); del r
        ts.append((time.perf_counter() - a) * 1e6)
    ms = ctx.timer_end(); wall = (time.perf_counter() - t0) * 1e3
    print(mode, "gpu ms/step %.3f wall ms/step %.3f host issue us:" % (ms / 8, wall / 8), [round(x) for x in ts])
    del keep
