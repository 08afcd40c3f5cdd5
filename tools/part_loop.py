"""Debug aid: the skewed-key partitioned case of tests/test_gpu_parity.py, repeated: python tools/part_loop.py [times] [hot fraction]"""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pyarrow as pa, kqgpu
from oracle import oracle as O
times = int(sys.argv[1]) if len(sys.argv) > 1 else 10
hot = float(sys.argv[2]) if len(sys.argv) > 2 else 0.9
rng = np.random.default_rng(78)
n = 1_500_000
keys_in = rng.integers(0, 400_000, n)
keys_in[rng.random(n) < hot] = 123456789
arrs = [pa.array(keys_in), pa.array(np.floor(rng.random(n) * 1000))]
ctx = kqgpu.Context(0); G = kqgpu.Engine(ctx)
def run(E, **kw):
    agg = E.HashAggregate([E.col(0)], [("SUM", E.col(1)), ("MIN", E.col(1)), ("MAX", E.col(1)), ("COUNT", E.col(1))], **kw)
    agg.update(E.RecordBatch.from_arrow(arrs))
    return {r[0]: r for r in zip(*[c.to_pylist() for c in agg.finalize().to_arrow()])}
want = run(O)
M = (1 << 64) - 1
def mix64(z):
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    return z ^ (z >> 31)
def part_of(key, log2=9):
    h = (0x9E3779B97F4A7C15 + 0) & M
    h = (mix64(h ^ (key & M)) + 0xD1B54A32D192ED03) & M
    return mix64(h) >> (64 - log2)
bad = 0
for it in range(times):
    got = run(G, expected_groups=400_000)
    missing = [k for k in want if k not in got]
    wrong = [k for k in got if k in want and got[k] != want[k]]
    extra = [k for k in got if k not in want]
    cnt = sum(r[4] for r in got.values())
    if os.environ.get("KQ_PART_DROP_SPILL"):
        hp = part_of(123456789)
        missing = [k for k in missing if part_of(k) != hp]; wrong = [k for k in wrong if part_of(k) != hp]       # only the other partitions are expected to be right
    if missing or wrong or extra:
        bad += 1
        print(f"run {it}: {len(got)} groups (want {len(want)}), missing {len(missing)}, wrong {len(wrong)}, extra {len(extra)}, rows counted {cnt} of {n}")
        import collections as _c
        print("    hot key partition", part_of(123456789), " missing by partition", _c.Counter(part_of(k) for k in missing).most_common(5),
              " wrong by partition", _c.Counter(part_of(k) for k in wrong).most_common(8))
        extra_rows = sum(got[k][4] - want[k][4] for k in wrong if got[k][4] > want[k][4]); lost_rows = sum(want[k][4] - got[k][4] for k in wrong if got[k][4] < want[k][4])
        print("    rows gained by wrong keys", extra_rows, " rows lost by wrong keys", lost_rows, " rows of missing keys", sum(want[k][4] for k in missing))
        def hk(key):
            h = (0x9E3779B97F4A7C15 + 0) & M
            h = (mix64(h ^ (key & M)) + 0xD1B54A32D192ED03) & M
            return mix64(h)
        # pair each missing single-row key with a wrong key that gained exactly its value
        gain = {}
        for k in wrong:
            d = got[k][1] - want[k][1]
            if got[k][4] == want[k][4] + 1: gain.setdefault(d, []).append(k)
        shown = 0
        for k in missing:
            v = want[k][1]
            if want[k][4] == 1 and len(gain.get(v, [])) == 1:
                g = gain[v][0]
                print(f"    lost key {k} hash {hk(k):016x} part {part_of(k)}  ->  gained by key {g} hash {hk(g):016x} part {part_of(g)}  (home distance at 2^23 slots: {(hk(g) >> 41) - (hk(k) >> 41)})")
                shown += 1
                if shown >= 12: break
        first_row = {}
        for i, key in enumerate(keys_in.tolist()): first_row.setdefault(key, i)
        lost_idx = sorted(first_row[key] for key in missing if want[key][4] == 1)
        dbl_idx = sorted(first_row[key] for key in wrong if want[key][4] == 1 and got[key][4] == 2 and got[key][1] == 2 * want[key][1])
        print("    input rows lost (single-row keys):", lost_idx[:40])
        print("    input rows counted twice (single-row keys):", dbl_idx[:40])
        print("    hot key: got", got.get(123456789), "want", want.get(123456789))
        for kk in wrong[:3]: print("    wrong", got[kk], "want", want[kk])
        for kk in missing[:3]: print("    missing", want[kk])
print("bad runs:", bad, "of", times)
