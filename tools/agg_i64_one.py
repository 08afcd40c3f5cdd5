import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200")); sys.path.insert(0, ROOT)
import kqgpu
rows = 50_000_000; groups = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ctx = kqgpu.Context(0); E = kqgpu.Engine(ctx)
specs = [dict(kind=1, col_id=0, ilo=0, ihi=groups), dict(kind=2, col_id=1, flo=0.0, fhi=1000.0)]
batch = E.generate(specs, 42, 0, rows); ctx.sync()
for _ in range(3):
    v = E.col(1)
    a = E.HashAggregate([E.col(0)], [("SUM", v), ("MIN", v), ("MAX", v), ("COUNT", v)], expected_groups=groups)
    ctx.timer_begin(); a.update(batch); r = a.finalize(); ms = ctx.timer_end()
print("i64", groups, rows, ms)
