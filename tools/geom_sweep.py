"""Sweep tile geometries of the aggregate kernel (KQ_AGG_GEOM = rows,warps[,share]) on one workload; every result is
compared with the first geometry's (keys/counts/extremes exact, Float64 sums within 1e-9).
   python tools/geom_sweep.py cfg3 100000000 10,7,1 6,14,2 ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200")); sys.path.insert(0, ROOT)
import kqgpu, bench
name = sys.argv[1]; rows = int(sys.argv[2]); geoms = sys.argv[3:]
ctx = kqgpu.Context(0); E = kqgpu.Engine(ctx)
wl = bench.WORKLOADS[name](rows)
batch = E.generate(wl.specs(), 42, 0, rows); ctx.sync()
base = None
keep = set(os.environ)
wl_nkeys = 2 if name == 'cfg5' else 1
for g in geoms:
    geom, *envs = g.split("@")           # "rows,warps,share@KQ_L2_PREFETCH=3@..." : extra tuning variables for this run
    for k in list(os.environ):
        if k.startswith("KQ_") and k not in keep: del os.environ[k]
    if geom != "default": os.environ["KQ_AGG_GEOM"] = geom
    for e in envs:
        k, v = e.split("="); os.environ[k] = v
    try:
        best = 1e9
        for _ in range(5):
            ctx.timer_begin(); r = wl.run(E, batch); ms = ctx.timer_end(); best = min(best, ms)
        got = sorted(zip(*[a.to_pylist() for a in r.to_arrow()]), key=lambda t: tuple((x is None, x) for x in t[:wl_nkeys]))
        ok = "base"
        if base is None: base = got
        else:
            ok = "same" if len(got) == len(base) else "ROWS DIFFER"
            for a, b in zip(got, base):
                for x, y in zip(a, b):
                    if x == y: continue
                    if isinstance(x, float) and isinstance(y, float) and abs(x - y) <= 1e-9 * max(abs(x), abs(y)): continue
                    ok = f"DIFF {a} vs {b}"
        print(f"{name} rows={rows} geom={g}: best {best:.3f} ms  {wl.algo_bytes(rows, len(got)) / best / 1e6:.0f} GB/s  [{ok}]", flush=True)
    except Exception as e:
        print(f"{name} geom={g}: FAILED {e}", flush=True)
