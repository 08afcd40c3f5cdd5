"""One CSV scan workload for profilers:  python tools/csv_one.py <records>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200")); sys.path.insert(0, ROOT)
import kqgpu, bench
rows = int(sys.argv[1])
ctx = kqgpu.Context(0); E = kqgpu.Engine(ctx)
wl = bench.WORKLOADS["csv"](rows)
dev = wl.prepare(E, ctx, 0, rows); ctx.sync()
for _ in range(3):
    ctx.timer_begin(); r = wl.run(E, dev); ms = ctx.timer_end(); n = r.row_count(); del r
print("csv", rows, n, ms)
