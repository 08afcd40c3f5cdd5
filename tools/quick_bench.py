"""Phase timing of single operators with the ctx timer (device time on the kernel stream)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200")); sys.path.insert(0, ROOT)
import kqgpu
import bench

def timeit(ctx, fn, reps=5):
    fn(); ctx.sync()
    best = 1e9
    for _ in range(reps):
        ctx.timer_begin(); r = fn(); ms = ctx.timer_end(); best = min(best, ms); del r
    return best

def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["cfg2f", "cfg2i", "proj", "cfg3", "cfg4", "cfg5"]
    ctx = kqgpu.Context(0); E = kqgpu.Engine(ctx)
    for name in which:
        if name == "proj":
            wl = bench.WORKLOADS["cfg2f"](rows)
            batch = E.generate(wl.specs(), 42, 0, rows); ctx.sync()
            proj = E.binary("ADD", E.binary("MUL", E.col(0), E.col(1)), E.col(2))
            ms = timeit(ctx, lambda: E.project([proj], batch))
            print(f"project a*b+c f64   rows={rows} {ms:8.3f} ms  {32*rows/ms/1e6:8.1f} GB/s  {rows/ms/1e6:8.2f} Grows/s", flush=True)
            continue
        wl = bench.WORKLOADS[name](rows)
        batch = E.generate(wl.specs(), 42, 0, rows); ctx.sync()
        t0 = time.perf_counter(); res = wl.run(E, batch); n_out = wl.result_rows(res); host_ms = (time.perf_counter() - t0) * 1e3
        ms = timeit(ctx, lambda: wl.run(E, batch))
        gb = wl.algo_bytes(rows, n_out) / ms / 1e6
        print(f"{name:6s} rows={rows} out={n_out} {ms:8.3f} ms  {gb:8.1f} GB/s  {rows/ms/1e6:8.2f} Grows/s  (first call wall {host_ms:.1f} ms)", flush=True)
        del batch, res

main()
