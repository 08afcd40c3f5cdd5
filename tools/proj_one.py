import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200")); sys.path.insert(0, ROOT)
import kqgpu, bench
rows = int(sys.argv[1])
ctx = kqgpu.Context(0); E = kqgpu.Engine(ctx)
wl = bench.WORKLOADS["cfg2f"](rows)
batch = E.generate(wl.specs(), 42, 0, rows); ctx.sync()
proj = E.binary("ADD", E.binary("MUL", E.col(0), E.col(1)), E.col(2))
for _ in range(4):
    ctx.timer_begin(); r = E.project([proj], batch); ms = ctx.timer_end(); del r
print("project", rows, ms)
