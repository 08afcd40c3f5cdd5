import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200")); sys.path.insert(0, ROOT)
import kqgpu, bench
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
ctx = kqgpu.Context(0); E = kqgpu.Engine(ctx)
wl = bench.WORKLOADS["cfg2f"](rows)
batch = E.generate(wl.specs(), 42, 0, rows); ctx.sync()
def t(fn, reps=5):
    fn(); ctx.sync(); best = 1e9
    for _ in range(reps):
        ctx.timer_begin(); r = fn(); ms = ctx.timer_end(); best = min(best, ms); del r
    return best
proj = E.binary("ADD", E.binary("MUL", E.col(0), E.col(1)), E.col(2))
for name, ka, kb in [("sel=0", 2.0, -1.0), ("sel=.25", 0.5, 0.5), ("sel=1", -1.0, 2.0)]:
    pred = E.binary("AND", E.binary("GT", E.col(0), E.lit_f64(ka)), E.binary("LT", E.col(1), E.lit_f64(kb)))
    ms = t(lambda: E.filter_project(pred, [proj], batch))
    ms2 = t(lambda: E.filter_project(pred, [], batch))
    ms3 = t(lambda: E.filter_project(E.binary("GT", E.col(0), E.lit_f64(ka)), [E.col(0)], batch))
    print(f"{name:8s} pred+proj {ms:7.3f} ms | pred only (no outputs) {ms2:7.3f} ms | a>k -> a {ms3:7.3f} ms", flush=True)
print("project only", t(lambda: E.project([proj], batch)))
