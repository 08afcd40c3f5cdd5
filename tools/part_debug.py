"""Partitioned vs plain aggregate path on the same table: per-column mismatch counts + timings."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200")); sys.path.insert(0, ROOT)
import kqgpu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
groups = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
ctx = kqgpu.Context(0); G = kqgpu.Engine(ctx)
specs = [dict(kind=1, col_id=0, ilo=0, ihi=groups), dict(kind=3, col_id=1, ilo=0, ihi=1000)]
batch = G.generate(specs, 42, 0, n); ctx.sync()
v = G.col(1)
aggs = [("SUM", v), ("MIN", v), ("MAX", v), ("COUNT", v)]
def run():
    best = 1e9
    for _ in range(3):
        ctx.timer_begin()
        a = G.HashAggregate([G.col(0)], aggs, expected_groups=groups)
        a.update(batch)
        r = a.finalize()
        best = min(best, ctx.timer_end())
    t = r.to_arrow()
    order = np.argsort(t[0].to_numpy(), kind="stable")
    return best, [c.to_numpy(zero_copy_only=False)[order] for c in t]
tp, part = run()
os.environ["KQ_NO_PARTITION"] = "1"
tq, plain = run()
if os.environ.get("KQ_TRY_UNBATCHED"):
    os.environ["KQ_GLOBAL_BATCHED"] = "0"
    tu, unb = run()
    print(f"unbatched plain {tu:.3f} ms groups {len(unb[0])}")
    for i, (a, b) in enumerate(zip(unb, plain)):
        if len(a) == len(b): print("  unbatched vs batched col", i, "mismatches", int((a != b).sum()))
if len(sys.argv) > 3:
    keys = batch.field(0).to_arrow().to_numpy(); vals = batch.field(1).to_arrow().to_numpy()
    truth = np.unique(keys)
    print("truth groups", len(truth), "partitioned-only", len(np.setdiff1d(part[0], truth)), "missing", len(np.setdiff1d(truth, part[0])),
          "plain-only", len(np.setdiff1d(plain[0], truth)), "missing", len(np.setdiff1d(truth, plain[0])))
    order = np.argsort(keys, kind="stable"); ks = keys[order]; vs = vals[order]
    starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
    tmin = np.minimum.reduceat(vs, starts); tmax = np.maximum.reduceat(vs, starts); tsum = np.add.reduceat(vs, starts); tcnt = np.diff(np.r_[starts, len(ks)])
    for name, res in (("partitioned", part), ("plain", plain)):
        if len(res[0]) == len(truth):
            print(name, "vs numpy: sum", (res[1] != tsum).sum(), "min", (res[2] != tmin).sum(), "max", (res[3] != tmax).sum(), "count", (res[4] != tcnt).sum())
print(f"partitioned {tp:.3f} ms  plain {tq:.3f} ms  rows {n} groups {len(part[0])} / {len(plain[0])}")
if len(part[0]) != len(plain[0]):
    u, c = np.unique(part[0], return_counts=True)
    print("duplicate keys in partitioned:", (c > 1).sum(), "missing:", len(np.setdiff1d(plain[0], part[0])))
else:
    for i, (a, b) in enumerate(zip(part, plain)):
        bad = np.flatnonzero(a != b)
        print("col", i, "mismatches", len(bad), [(part[0][j], a[j], b[j]) for j in bad[:5]])
    print("count sum", part[4].sum(), plain[4].sum(), n)
