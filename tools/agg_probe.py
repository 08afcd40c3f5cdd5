import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200")); sys.path.insert(0, ROOT)
import kqgpu, bench
rows = 100_000_000
ctx = kqgpu.Context(0); E = kqgpu.Engine(ctx)
def t(name, specs, hint, aggs=("SUM","MIN","MAX","COUNT")):
    batch = E.generate(specs, 42, 0, rows); ctx.sync()
    def run():
        v = E.col(1)
        a = E.HashAggregate([E.col(0)], [(k, v) for k in aggs], expected_groups=hint)
        a.update(batch)
        return a.finalize()
    best = 1e9
    for _ in range(4):
        ctx.timer_begin(); r = run(); ms = ctx.timer_end(); best = min(best, ms); n = r.row_count(); del r
    print(f"{name:40s} out={n:6d} {best:8.3f} ms {rows/best/1e6:8.2f} Grows/s", flush=True)
V = dict(kind=2, col_id=1, flo=0.0, fhi=1000.0)
t("i64 key 40 groups hint 40", [dict(kind=1, col_id=0, ilo=0, ihi=40), V], 40)
t("i64 key 40 groups hint 0", [dict(kind=1, col_id=0, ilo=0, ihi=40), V], 0)
t("i64 key 40 groups offset 1000", [dict(kind=1, col_id=0, ilo=1000, ihi=1040), V], 40)
t("utf8 key 50 groups hint 50", [dict(kind=5, col_id=0, dict=bench.STATES, dict_width=2), V], 50)
t("i64 key 40 groups SUM only", [dict(kind=1, col_id=0, ilo=0, ihi=40), V], 40, ("SUM",))
t("i64 key 40 groups SUM,COUNT", [dict(kind=1, col_id=0, ilo=0, ihi=40), V], 40, ("SUM","COUNT"))
t("i64 key 40 groups MIN,MAX", [dict(kind=1, col_id=0, ilo=0, ihi=40), V], 40, ("MIN","MAX"))
