"""Aggregate throughput vs number of groups (Int64 key, SUM/MIN/MAX/COUNT(Float64)): where the global-table path stands."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200")); sys.path.insert(0, ROOT)
import kqgpu
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
ctx = kqgpu.Context(0); E = kqgpu.Engine(ctx)
for groups in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "40,1000,100000,1000000,10000000".split(","))]:
    specs = [dict(kind=1, col_id=0, ilo=0, ihi=groups), dict(kind=2, col_id=1, flo=0.0, fhi=1000.0)]
    batch = E.generate(specs, 42, 0, rows); ctx.sync()
    def run():
        v = E.col(1)
        a = E.HashAggregate([E.col(0)], [("SUM", v), ("MIN", v), ("MAX", v), ("COUNT", v)], expected_groups=groups)
        a.update(batch)
        return a.finalize()
    best = 1e9
    for _ in range(4):
        ctx.timer_begin(); r = run(); ms = ctx.timer_end(); best = min(best, ms); n = r.row_count(); del r
    print(f"groups={groups:9d} out={n:9d} {best:8.3f} ms {rows/best/1e6:8.2f} Grows/s {16*rows/best/1e6:8.1f} GB/s", flush=True)
    del batch
