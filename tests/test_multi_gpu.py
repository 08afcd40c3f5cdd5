"""Multi-GPU merge of partial aggregates (main()'s partition -> partial -> merge, Main.kt:1306-1342) over NCCL.

One rank per GPU, here as one thread + one kq_ctx per device inside a single process (the C ABI has no
global state; bench.py runs the same calls as one process per GPU under torchrun). Skipped on boxes
with a single GPU. The checker is the CPU oracle on the whole (unsharded) table.
"""
import ctypes as C
import threading

import pytest

pytestmark = pytest.mark.gpu

STATES = "ALAKAZARCACOCTDEFLGAHIIDILINIAKSKYLAMEMDMAMIMNMSMOMTNENVNHNJNMNYNCNDOHOKORPARISCSDTNTXUTVTVAWAWVWIWY"


def run_ranks(gpu, world, body):
    """body(rank, ctx, engine) on `world` threads, each with its own context + NCCL communicator."""
    L = gpu.lib()
    ctxs = [gpu.Context(r) for r in range(world)]
    idbuf = C.create_string_buffer(gpu.COMM_ID_BYTES)
    ctxs[0].check(L.kq_comm_unique_id(ctxs[0].h, idbuf))
    out, errs = [None] * world, []

    def work(r):
        try:
            ctx = ctxs[r]
            ctx.check(L.kq_comm_init(ctx.h, idbuf.raw, r, world))
            out[r] = body(r, ctx, gpu.Engine(ctx))
            ctx.check(L.kq_comm_barrier(ctx.h))
            ctx.check(L.kq_comm_destroy(ctx.h))
        except Exception as e:   # noqa: BLE001
            errs.append((r, e))

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join(300) for t in th]
    assert not errs, errs
    return out


def rows_of(batch):
    cols = [a.to_pylist() for a in batch.to_arrow()]
    return sorted(zip(*cols), key=lambda t: tuple((x is None, x) for x in t[:1]))


def close(a, b, rtol=1e-9):
    assert len(a) == len(b), (len(a), len(b))
    for ra, rb in zip(a, b):
        assert len(ra) == len(rb)
        for x, y in zip(ra, rb):
            if isinstance(x, float) and isinstance(y, float):
                assert abs(x - y) <= rtol * max(abs(x), abs(y), 1e-300), (ra, rb)
            else:
                assert x == y, (ra, rb)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_low_cardinality_allreduce_merge(gpu, oracle, world):
    if gpu.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    n = 300_000
    specs = [dict(kind=5, col_id=0, dict=STATES, dict_width=2, null_per_10k=300), dict(kind=2, col_id=1, flo=0.0, fhi=1000.0, null_per_10k=500),
             dict(kind=3, col_id=2, ilo=0, ihi=1000)]

    def plan(E, batch):
        a = E.HashAggregate([E.col(0)], [("SUM", E.col(1)), ("MIN", E.col(1)), ("MAX", E.col(1)), ("COUNT", E.col(1)), ("SUM", E.col(2))])
        a.update(batch)
        return a

    def body(r, ctx, E):
        lo, hi = r * n // world, (r + 1) * n // world
        a = plan(E, E.generate(specs, 7, lo, hi))
        a.merge_allreduce()
        return rows_of(a.finalize())

    got = run_ranks(gpu, world, body)
    want = rows_of(plan(oracle, oracle.generate(specs, 7, 0, n)).finalize())
    for r in range(world):          # every rank holds the full merged result
        close(got[r], want)
    assert all(g == got[0] for g in got), "all-reduce must leave bit-identical results on every rank"


@pytest.mark.parametrize("world", [2, 4])
def test_mid_cardinality_allreduce_merge(gpu, oracle, world):
    """More partial groups per rank than the one-shot small merge carries (1024): the union-dictionary + dense
    ncclAllReduce path of kq_hashagg_merge_allreduce."""
    if gpu.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    n = 200_000
    specs = [dict(kind=3, col_id=0, ilo=0, ihi=3000), dict(kind=2, col_id=1, flo=-5.0, fhi=5.0, null_per_10k=500)]

    def plan(E, batch):
        a = E.HashAggregate([E.col(0)], [("SUM", E.col(1)), ("MIN", E.col(1)), ("MAX", E.col(1)), ("COUNT", E.col(1))])
        a.update(batch)
        return a

    def body(r, ctx, E):
        lo, hi = r * n // world, (r + 1) * n // world
        a = plan(E, E.generate(specs, 11, lo, hi))
        a.merge_allreduce()
        return rows_of(a.finalize())

    got = run_ranks(gpu, world, body)
    want = rows_of(plan(oracle, oracle.generate(specs, 11, 0, n)).finalize())
    for r in range(world):
        close(got[r], want)
    assert all(g == got[0] for g in got), "all-reduce must leave bit-identical results on every rank"


@pytest.mark.parametrize("world", [2, 4, 8])
def test_high_cardinality_alltoall_repartition(gpu, oracle, world):
    if gpu.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    n = 400_000
    specs = [dict(kind=1, col_id=0, ilo=0, ihi=50_000), dict(kind=2, col_id=1, flo=0.0, fhi=1000.0)]

    def plan(E, batch, **kw):
        v = E.col(1)
        a = E.HashAggregate([E.col(0)], [("SUM", v), ("MIN", v), ("MAX", v), ("COUNT", v)], **kw)
        a.update(batch)
        return a

    def body(r, ctx, E):
        lo, hi = r * n // world, (r + 1) * n // world
        a = plan(E, E.generate(specs, 11, lo, hi), expected_groups=50_000)
        a.repartition_alltoall()
        return rows_of(a.finalize())

    got = run_ranks(gpu, world, body)
    want = rows_of(plan(oracle, oracle.generate(specs, 11, 0, n)).finalize())
    keys = [set(t[0] for t in g) for g in got]
    for i in range(world):
        for j in range(i + 1, world):
            assert not (keys[i] & keys[j]), "after the repartition every key lives on exactly one rank"
    merged = sorted((t for g in got for t in g), key=lambda t: t[0])
    close(merged, want)


@pytest.mark.parametrize("world", [2, 4])
def test_long_utf8_keys_travel_with_the_merge(gpu, oracle, world):
    """Utf8 group keys longer than 7 bytes are interned per rank (key heap); the all-gather merge ships the strings along, so
    every rank can emit every group — also groups none of whose rows it saw itself."""
    if gpu.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import numpy as np
    import pyarrow as pa
    n = 40_000
    rng = np.random.default_rng(5)
    names = np.array(["Stockholm", "Uppsala", "Sthlm", "Göteborg by the sea", "Örnsköldsvik", "rank-private-key-%d"], dtype=object)
    k = names[rng.integers(0, 5, n)].copy()
    for r in range(world):                       # one long key that only rank r ever sees
        lo, hi = r * n // world, (r + 1) * n // world
        k[lo:lo + 50] = "rank-private-key-%d" % r
    v = np.floor(rng.random(n) * 100)
    arrs = [pa.array(k, pa.string()), pa.array(v, pa.float64())]

    def plan(E, a0, a1):
        a = E.HashAggregate([E.col(0)], [("SUM", E.col(1)), ("MIN", E.col(1)), ("MAX", E.col(1)), ("COUNT", E.col(1))], **({"expected_groups": 16} if E is not oracle else {}))
        a.update(E.RecordBatch.from_arrow([a0, a1]))
        return a

    def body(r, ctx, E):
        lo, hi = r * n // world, (r + 1) * n // world
        a = plan(E, arrs[0].slice(lo, hi - lo), arrs[1].slice(lo, hi - lo))
        a.merge_allreduce()
        return rows_of(a.finalize())

    got = run_ranks(gpu, world, body)
    want = rows_of(plan(oracle, arrs[0], arrs[1]).finalize())
    for r in range(world):
        close(got[r], want)
    assert all(g == got[0] for g in got)


@pytest.mark.parametrize("world", [2])
def test_sql_query_with_the_merge_inside_the_aggregate(gpu, oracle, world):
    """main() (Main.kt:1306-1342) at the level of the reference's API: every rank runs the same SQL over its shard with
    ExecutionContext(engine, merge="allreduce"); the aggregate merges the partial tables collectively, so every rank's single
    output batch is the answer over the whole table (checked against the oracle running the query on the unsharded table)."""
    if gpu.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from kqgpu import plan as P
    n = 200_000
    specs = [dict(kind=5, col_id=0, dict=STATES, dict_width=2), dict(kind=2, col_id=1, flo=0.0, fhi=1000.0)]
    sql = "SELECT state, SUM(v), MIN(v), MAX(v), COUNT(v) FROM t WHERE v >= 10.0 GROUP BY state"
    schema = P.Schema([P.Field("state", P.StringType), P.Field("v", P.DoubleType)])

    def query(E, lo, hi, merge):
        q = P.ExecutionContext(E, merge=merge)
        mid = (lo + hi) // 2
        q.registerDataSource("t", P.InMemoryDataSource(E, schema, [E.generate(specs, 11, lo, mid), E.generate(specs, 11, mid, hi)]))
        out = list(q.execute(q.sql(sql)))
        assert len(out) == 1
        return rows_of(out[0])

    got = run_ranks(gpu, world, lambda r, ctx, E: query(E, r * n // world, (r + 1) * n // world, "allreduce"))
    want = query(oracle, 0, n, None)
    assert len(want) == 50
    for r in range(world):
        close(got[r], want)
