import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200")); sys.path.insert(0, ROOT)
import kqgpu, bench
name = sys.argv[1]; rows = int(sys.argv[2])
ctx = kqgpu.Context(0); E = kqgpu.Engine(ctx)
wl = bench.WORKLOADS[name](rows)
batch = E.generate(wl.specs(), 42, 0, rows); ctx.sync()
for _ in range(4):
    ctx.timer_begin(); r = wl.run(E, batch); ms = ctx.timer_end(); del r
print(name, rows, ms)
