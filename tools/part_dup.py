"""Debug aid: the partitioned-path parity case of tests/test_gpu_parity.py, printing the groups that differ from the oracle.
   python tools/part_dup.py [null_frac] [one|two batches]"""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pyarrow as pa, kqgpu
from oracle import oracle as O
from test_gpu_parity import rand_table, AGGS
null_frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
two = (sys.argv[2] if len(sys.argv) > 2 else "two") == "two"
rng = np.random.default_rng(77)
n = 1_100_000
arrs = rand_table(rng, n, null_frac)
ctx = kqgpu.Context(0); G = kqgpu.Engine(ctx)
def run(E, **kw):
    agg = E.HashAggregate([E.col(0)], [(kind, E.col(c)) for kind, c in AGGS], **kw)
    agg.update(E.RecordBatch.from_arrow(arrs))
    if two: agg.update(E.RecordBatch.from_arrow([a.slice(7, 1_050_000) for a in arrs]))
    out = agg.finalize()
    return list(zip(*[c.to_pylist() for c in out.to_arrow()]))
got, want = run(G, expected_groups=700_000), run(O)
print("rows", len(got), len(want))
cg = collections.Counter(r[0] for r in got); cw = collections.Counter(r[0] for r in want)
dups = [k for k, v in cg.items() if v > 1]
print("keys more than once on the GPU:", len(dups), dups[:10])
print("keys only on the GPU:", [k for k in cg if k not in cw][:10], " only in the oracle:", [k for k in cw if k not in cg][:10])
byk = collections.defaultdict(list)
for r in got: byk[r[0]].append(r)
wk = {r[0]: r for r in want}
for k in dups[:6]:
    print("key", k); [print("   gpu   ", r) for r in byk[k]]; print("   oracle", wk.get(k))
