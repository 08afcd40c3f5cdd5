"""Per-tile pipeline timestamps of the filter kernel (KQ_TRACE build): python tools/trace_fp.py rows out.npy"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200")); sys.path.insert(0, ROOT)
rows = int(sys.argv[1]); out = sys.argv[2]
os.environ["KQ_TRACE_FILE"] = out
os.environ["KQ_JIT_CACHE"] = "off"
import numpy as np
import kqgpu, bench
ctx = kqgpu.Context(0); E = kqgpu.Engine(ctx)
wl = bench.WORKLOADS["cfg2f"](rows)
batch = E.generate(wl.specs(), 42, 0, rows); ctx.sync()
for _ in range(3):
    r = wl.run(E, batch); ctx.sync(); del r
t = np.fromfile(out, dtype=np.uint64).reshape(-1, 8).astype(np.int64)
t0 = t[:, 0].min()
t = np.where(t > 0, t - t0, -1)
names = ["issue", "data", "agg", "lb_start", "lb_done", "B_wait", "B_go"]
print("tiles", len(t), "span us", t.max() / 1e3)
def stat(name, a):
    print(f"{name:28s} mean {a.mean()/1e3:7.2f}  p50 {np.percentile(a,50)/1e3:7.2f}  p90 {np.percentile(a,90)/1e3:7.2f}  max {a.max()/1e3:7.2f} us")
mid = t[len(t)//4: 3*len(t)//4]
stat("issue->data (load)", mid[:,1]-mid[:,0])
stat("data->agg (step A)", mid[:,2]-mid[:,1])
stat("agg->lb_start", mid[:,3]-mid[:,2])
stat("lb_start->lb_done", mid[:,4]-mid[:,3])
stat("agg->lb_done", mid[:,4]-mid[:,2])
stat("B_wait->B_go (stall)", mid[:,6]-mid[:,5])
stat("issue->B_go (total)", mid[:,6]-mid[:,0])
stat("agg(t)-agg(t-1) skew", np.abs(np.diff(mid[:,2])))
d = mid[1:,2] - np.maximum.accumulate(mid[:,2])[:-1]
stat("agg(t) - max agg(<t)", -d)
np.save(out + ".npy", t)
