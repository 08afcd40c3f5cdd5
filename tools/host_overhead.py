"""Host-side cost of one operator call (tiny input: the GPU work is negligible)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200")); sys.path.insert(0, ROOT)
import kqgpu, bench
ctx = kqgpu.Context(0); E = kqgpu.Engine(ctx)
wl = bench.WORKLOADS["cfg2f"](10000)
batch = E.generate(wl.specs(), 42, 0, 10000); ctx.sync()
pred, proj = wl.exprs(E)
for name, fn in [("filter_project (exprs prebuilt)", lambda: E.filter_project(pred, proj, batch)), ("wl.run (exprs rebuilt)", lambda: wl.run(E, batch)),
                 ("project", lambda: E.project(proj, batch))]:
    for _ in range(3): r = fn(); del r
    ctx.sync(); t0 = time.perf_counter()
    for _ in range(200): r = fn(); del r
    t1 = time.perf_counter(); ctx.sync(); t2 = time.perf_counter()
    print(f"{name}: {(t1 - t0) / 200 * 1e6:.1f} us/call issue, {(t2 - t0) / 200 * 1e6:.1f} us/call incl. drain")
