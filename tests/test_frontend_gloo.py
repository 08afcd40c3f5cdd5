"""main() of the reference (Main.kt:1306-1342) across processes, at the level of its own API: every rank of a world_size-2
`gloo` group runs the partial query over its shard of the rows through ExecutionContext.sql, the partial batches are
collected on every rank (main() collects them from its coroutines), registered as an in-memory table, and the merge query
runs over them. The result must equal the same query over the whole table. On GPUs the collection and the merge query
are one collective inside the aggregate (ExecutionContext(engine, merge="allreduce") -> kq_hashagg_merge_allreduce,
tests/test_multi_gpu.py); here the engine is the CPU oracle."""
import os
import sys

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STATES = "ALAKAZARCACOCTDEFLGAHIIDILINIAKSKYLAMEMDMAMIMNMSMOMTNENVNHNJNMNYNCNDOHOKORPARISCSDTNTXUTVTVAWAWVWIWY"
SPECS = [dict(kind=5, col_id=0, dict=STATES, dict_width=2), dict(kind=1, col_id=1, ilo=0, ihi=1000)]      # Utf8 key, Int64 values (generator kind 1): sums are exact
N_PER_RANK = 30_000
PARTIAL = "SELECT state, MAX(v) AS max_v, MIN(v) AS min_v, SUM(v) AS sum_v, COUNT(v) AS n FROM t GROUP BY state"
MERGE = "SELECT state, MAX(max_v), MIN(min_v), SUM(sum_v), SUM(n) FROM partials GROUP BY state ORDER BY state"


def setup_paths():
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "query-engines_b200")]


def run_partial(O, P, lo, hi, batches=3):
    step = (hi - lo) // batches
    data = [O.generate(SPECS, 3, lo + i * step, lo + (i + 1) * step if i < batches - 1 else hi) for i in range(batches)]
    ctx = P.ExecutionContext(O)
    ctx.registerDataSource("t", P.InMemoryDataSource(O, P.Schema([P.Field("state", P.StringType), P.Field("v", P.Int64Type)]), data))
    df = ctx.sql(PARTIAL)
    return df.schema(), [[a.to_pylist() for a in b.to_arrow()] for b in ctx.execute(df)]


def run_merge(O, P, schema, partial_tables):
    import pyarrow as pa
    types = {P.StringType: pa.string(), P.Int64Type: pa.int64()}
    batches = [O.RecordBatch.from_arrow([pa.array(col, type=types[f.dataType]) for col, f in zip(cols, schema.fields)]) for cols in partial_tables]
    ctx = P.ExecutionContext(O)
    ctx.registerDataSource("partials", P.InMemoryDataSource(O, schema, batches))
    out = ctx.execute(ctx.sql(MERGE))
    return sorted(r for b in out for r in zip(*[a.to_pylist() for a in b.to_arrow()]))


def worker(rank, world, port, q):
    setup_paths()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from kqgpu import plan as P
    from oracle import oracle as O
    O.build()
    lo, hi = bench.shard_range(rank, N_PER_RANK)
    schema, mine = run_partial(O, P, lo, hi)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)                         # every rank receives every rank's partial batches
    q.put((rank, run_merge(O, P, schema, [t for part in gathered for t in part])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_partition_partial_merge_through_the_sql_api_equals_the_whole_table():
    world = 2
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [mpctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=240) for _ in range(world))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    setup_paths()
    from kqgpu import plan as P
    from oracle import oracle as O
    O.build()
    schema, whole = run_partial(O, P, 0, world * N_PER_RANK, batches=1)
    want = sorted(zip(*whole[0]))
    assert len(want) == 50
    for rank, merged in res:
        assert merged == want, f"rank {rank}: merged partials differ from the whole-table query"
