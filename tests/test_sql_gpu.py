"""The reference's SQL / DataFrame front-end (kqgpu/plan.py, kqgpu/sql.py) driving the GPU operators: the same query strings
through ExecutionContext.sql on kqgpu.Engine and on the CPU oracle must agree, and the reference's own queries on its
fixture must give the golden vectors. (The front-end itself is checked on the CPU in tests/test_frontend.py; this file
sorts last on purpose.)"""
import json
import os

import pytest

from kqgpu import plan as P

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "employee_golden.json")


@pytest.fixture(scope="module")
def G(gpu, gctx):
    return gpu.Engine(gctx)


def rows(batches):
    out = []
    for b in batches:
        out += list(zip(*[a.to_pylist() for a in b.to_arrow()]))
    return out


def test_reference_main_on_the_fixture(G):
    """main() (Main.kt:1306-1342): partial aggregates per partition from the CSV source, merged by a second query over an
    in-memory table of the partial batches."""
    g = json.load(open(GOLDEN, encoding="utf-8"))
    text = bytes.fromhex(g["csv_text_hex"])
    partials, df = [], None
    for _ in range(3):
        ctx = P.ExecutionContext(G)
        ctx.registerDataSource("tripdata", P.CsvDataSource(G, text, True))
        df = ctx.sql("SELECT state, MAX(CAST(salary AS double)) AS max_amount FROM tripdata GROUP BY state")
        partials += list(ctx.execute(df))
    assert len(partials) == 3 and dict(rows(partials[:1])) == g["group_by_state_max_salary"]
    ctx = P.ExecutionContext(G)
    ctx.registerDataSource("tripdata", P.InMemoryDataSource(G, df.schema(), partials))
    out = ctx.execute(ctx.sql("SELECT state, MAX(max_amount) FROM tripdata GROUP BY state ORDER BY max_amount"))
    assert dict(rows(out)) == g["group_by_state_max_salary"]


@pytest.mark.parametrize("state,key", [("CO", "config1_where_state_eq_CO"), ("Uppsala", "where_state_eq_Uppsala")])
def test_baseline_config1_query_string(G, state, key):
    g = json.load(open(GOLDEN, encoding="utf-8"))
    ctx = P.ExecutionContext(G)
    ctx.registerDataSource("employee", P.CsvDataSource(G, bytes.fromhex(g["csv_text_hex"]), True))
    names = ["id", "first_name", "last_name", "state", "salary"]
    batches = list(ctx.execute(ctx.sql(f"SELECT id, first_name, last_name, state, salary FROM employee WHERE state = '{state}'")))
    got = {n: [] for n in names}
    for b in batches:
        for n, a in zip(names, b.to_arrow()):
            got[n] += a.to_pylist()
    assert got == g[key]


def test_baseline_config2_and_config3_query_strings_agree_with_the_oracle(G, oracle):
    specs2 = [dict(kind=2, col_id=i, flo=0.0, fhi=1.0) for i in range(3)]
    specs3 = [dict(kind=5, col_id=0, dict="ALAKAZARCACOCTDE", dict_width=2), dict(kind=2, col_id=1, flo=0.0, fhi=1000.0)]
    res = {}
    for name, E in (("gpu", G), ("oracle", oracle)):
        ctx = P.ExecutionContext(E)
        t2 = [E.generate(specs2, 42, i * 50_000, (i + 1) * 50_000) for i in range(4)]
        t3 = [E.generate(specs3, 42, i * 50_000, (i + 1) * 50_000) for i in range(4)]
        ctx.registerDataSource("t2", P.InMemoryDataSource(E, P.Schema([P.Field(c, P.DoubleType) for c in "abc"]), t2))
        ctx.registerDataSource("t3", P.InMemoryDataSource(E, P.Schema([P.Field("state", P.StringType), P.Field("v", P.DoubleType)]), t3))
        q2 = rows(ctx.execute(ctx.sql("SELECT a * b + c AS r FROM t2 WHERE a > 0.5 AND b < 0.5")))
        q3 = sorted(rows(ctx.execute(ctx.sql("SELECT state, SUM(v), MIN(v), MAX(v), COUNT(v) FROM t3 GROUP BY state"))))
        res[name] = (q2, q3)
    assert res["gpu"][0] == res["oracle"][0] and len(res["gpu"][0]) > 0                     # Float64 expressions: bit-exact (no FMA)
    assert len(res["gpu"][1]) == len(res["oracle"][1]) == 8
    for a, b in zip(res["gpu"][1], res["oracle"][1]):
        assert a[0] == b[0] and a[2:] == b[2:] and abs(a[1] - b[1]) <= 1e-9 * abs(b[1])       # SUM within 1e-9 relative
