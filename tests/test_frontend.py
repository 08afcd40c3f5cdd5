"""The reference's DataFrame / logical-plan API, planner and SQL front-end (kqgpu/plan.py, kqgpu/sql.py; Main.kt:56-175,
359-446, 662-770, 807-1290) — checked without a device by running the plans on the CPU oracle, which speaks the same
operator vocabulary as kqgpu.Engine: the reference's own queries (main(), Main.kt:1306-1342) on its only fixture against the
golden vectors, its plan printing and error behaviour, and the extensions that BASELINE.json's query strings need (WHERE,
literals, binary operators, SUM/MIN/COUNT) against direct operator calls."""
import json
import os

import pytest

from kqgpu import plan as P
from kqgpu import sql as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "employee_golden.json")


@pytest.fixture(scope="module")
def golden():
    return json.load(open(GOLDEN, encoding="utf-8"))


@pytest.fixture()
def ctx(oracle, golden):
    c = P.ExecutionContext(oracle)
    c.registerDataSource("employee", P.CsvDataSource(oracle, bytes.fromhex(golden["csv_text_hex"]), True))
    return c


def rows(batches):
    out = []
    for b in batches:
        out += list(zip(*[a.to_pylist() for a in b.to_arrow()])) if b.num_columns() else []
    return out


def table(batches, names):
    cols = {n: [] for n in names}
    for b in batches:
        for n, a in zip(names, b.to_arrow()):
            cols[n] += a.to_pylist()
    return cols


# ---------------------------------------------------------------- tokenizer and parser
def test_tokenizer_token_kinds():
    toks = S.SqlTokenizer("SELECT a1, `odd name`, MAX(x) AS m FROM t WHERE y >= 1.5 AND z <> 'it''s'").tokenize().tokens
    kinds = [(t.text, t.type) for t in toks]
    assert kinds[:4] == [("SELECT", "KEYWORD"), ("a1", "IDENTIFIER"), (",", "SYMBOL"), ("odd name", "IDENTIFIER")]
    assert ("MAX", "KEYWORD") in kinds and (">=", "SYMBOL") in kinds and ("1.5", "DOUBLE") in kinds and ("<>", "SYMBOL") in kinds
    assert kinds[-2:] == [("it", "STRING"), ("s", "STRING")]          # the reference's strings end at the next quote (Main.kt:1031-1034)
    with pytest.raises(S.TokenizeException):
        S.SqlTokenizer("SELECT 'open").tokenize()


def test_parser_builds_the_reference_tree():
    ast = S.SqlParser(S.SqlTokenizer("SELECT VendorID, MAX(CAST(fare_amount AS double)) AS max_amount FROM tripdata GROUP BY VendorID").tokenize()).parse()
    assert ast == S.SqlSelect([S.SqlIdentifier("VendorID"),
                               S.SqlAlias(S.SqlFunction("MAX", [S.SqlCast(S.SqlIdentifier("fare_amount"), S.SqlIdentifier("double"))]), S.SqlIdentifier("max_amount"))],
                              None, [S.SqlIdentifier("VendorID")], [], "tripdata")
    ast = S.SqlParser(S.SqlTokenizer("SELECT a FROM t ORDER BY a, b DESC").tokenize()).parse()
    assert ast.orderBy == [S.SqlSort(S.SqlIdentifier("a"), True), S.SqlSort(S.SqlIdentifier("b"), False)]


def test_operator_precedence_and_parentheses():
    def where(sql):
        return S.SqlParser(S.SqlTokenizer(f"SELECT x FROM t WHERE {sql}").tokenize()).parse().selection
    B, I, L = S.SqlBinary, S.SqlIdentifier, S.SqlLiteral
    assert where("a > 1 AND b < 2 OR c = 3") == B("OR", B("AND", B("GT", I("a"), L(1)), B("LT", I("b"), L(2))), B("EQ", I("c"), L(3)))
    assert where("a * b + c") == B("ADD", B("MUL", I("a"), I("b")), I("c"))
    assert where("a * (b + c)") == B("MUL", I("a"), B("ADD", I("b"), I("c")))
    assert where("a - b - c") == B("SUB", B("SUB", I("a"), I("b")), I("c"))              # left-associative
    assert where("a > -0.5") == B("GT", I("a"), L(-0.5))


@pytest.mark.parametrize("sql,exc,match", [
    ("SELECT a b FROM t", P.IllegalStateException, "Expected FROM"),
    ("SELECT a FROM t GROUP a", P.IllegalStateException, "Unexpected token"),
    ("SELECT FROM t", P.IllegalStateException, "Unexpected token"),
    ("SELECT a FROM (t)", P.SQLException, "No table named"),
    ("SELECT CAST(a) FROM t", P.SQLException, "CAST"),
])
def test_parser_errors(ctx, sql, exc, match):
    with pytest.raises(exc, match=match):
        ctx.sql(sql)


# ---------------------------------------------------------------- the reference's queries on its fixture
def test_schema_and_scan_of_the_fixture(ctx, golden):
    df = ctx.sql("SELECT id, first_name, last_name, state, job_title, salary FROM employee")
    assert df.schema().names() == golden["schema"] and all(f.dataType == P.StringType for f in df.schema().fields)
    got = table(ctx.execute(df), golden["schema"])
    assert got == golden["columns"]
    assert got["last_name"][2].encode("utf-8").hex() == golden["last_name_row3_utf8_hex"]


def test_main_query_plans_print_like_the_reference(ctx):
    df = ctx.sql("SELECT state, MAX(CAST(salary AS double)) AS max_amount FROM employee GROUP BY state")
    assert P.format_plan(df.logicalPlan()) == (
        "Projection: #0,#1 as max_amount\n"
        "\tAggregate: groupExpr=[#state], aggregateExpr=[MAX(CAST(#salary AS FloatingPoint(DOUBLE)))]\n"
        "\t\tScan: employee; projection=None\n")
    opt = P.ProjectionPushDownRule().optimize(df.logicalPlan())
    assert str(opt.children()[0].children()[0]) == "Scan: employee; projection=[salary, state]"       # pushed down, sorted (Main.kt:763-765)
    assert df.schema() == P.Schema([P.Field("state", P.StringType), P.Field("max_amount", P.DoubleType)])
    phys = P.createPhysicalPlan(opt, ctx.engine)
    assert [type(n).__name__ for n in (phys, phys.children()[0], phys.children()[0].children()[0])] == ["ProjectionExec", "HashAggregateExec", "ScanExec"]
    assert phys.children()[0].schema().names() == ["state", "MAX"]


def test_main_partition_partial_merge(ctx, oracle, golden):
    """main() (Main.kt:1306-1342): the partial query per partition, its batches registered as an in-memory table, the merge
    query over them — ORDER BY is parsed and ignored."""
    partial_sql = "SELECT state, MAX(CAST(salary AS double)) AS max_amount FROM employee GROUP BY state"
    partials = []
    for _ in range(3):                                             # three "months" of the same file
        df = ctx.sql(partial_sql)
        partials += list(ctx.execute(df))
    assert all(b.row_count() == 2 for b in partials) and len(partials) == 3
    merge = P.ExecutionContext(oracle)
    merge.registerDataSource("tripdata", P.InMemoryDataSource(oracle, df.schema(), partials))
    out = merge.sql("SELECT state, MAX(max_amount) FROM tripdata GROUP BY state ORDER BY max_amount")
    assert out.schema().names() == ["state", "MAX"]
    assert dict(rows(merge.execute(out))) == golden["group_by_state_max_salary"]
    assert dict(rows(ctx.execute(ctx.sql(partial_sql)))) == golden["group_by_state_max_salary"]


@pytest.mark.parametrize("state,key", [("CO", "config1_where_state_eq_CO"), ("Uppsala", "where_state_eq_Uppsala")])
def test_baseline_config1_by_sql_and_by_dataframe(ctx, oracle, golden, state, key):
    """BASELINE.json configs[0]: employee.csv via CsvDataSource: SELECT id, first_name, last_name, state, salary WHERE state = 'CO'."""
    names = ["id", "first_name", "last_name", "state", "salary"]
    df = ctx.sql(f"SELECT id, first_name, last_name, state, salary FROM employee WHERE state = '{state}'")
    assert table(ctx.execute(df), names) == golden[key]
    # the same through the DataFrame API (Main.kt:359-364, with filter as the extension)
    src = P.CsvDataSource(oracle, bytes.fromhex(golden["csv_text_hex"]), True)
    df2 = P.DataFrame(P.Scan("employee.csv", src, [])).filter(P.BinaryExpr("EQ", P.col("state"), P.lit(state))).project([P.col(n) for n in names])
    assert table(ctx.execute(df2), names) == golden[key]
    assert P.format_plan(df2.logicalPlan()) == (
        "Projection: #id,#first_name,#last_name,#state,#salary\n"
        f"\tSelection: #state = '{state}'\n"
        "\t\tScan: employee.csv; projection=None\n")
    # the selection is folded into the projection's kernel, and the scan reads only what the plan names
    phys = P.createPhysicalPlan(P.ProjectionPushDownRule().optimize(df2.logicalPlan()), oracle)
    assert type(phys).__name__ == "ProjectionExec" and phys.predicate is not None and type(phys.children()[0]).__name__ == "ScanExec"
    assert phys.children()[0].projection == sorted(names)


def test_csv_file_on_disk_and_missing_file(tmp_path, oracle, golden):
    path = tmp_path / "employee.csv"
    path.write_bytes(bytes.fromhex(golden["csv_text_hex"]))
    c = P.ExecutionContext(oracle)
    c.registerCsv("employee", str(path))
    assert dict(rows(c.execute(c.sql("SELECT state, MAX(CAST(salary AS double)) FROM employee GROUP BY state")))) == golden["group_by_state_max_salary"]
    with pytest.raises(FileNotFoundError):                         # Scan derives its schema when it is built (Main.kt:106, 306-309)
        c.registerCsv("nope", str(tmp_path / "nope.csv"))


@pytest.mark.parametrize("sql,exc,match", [
    ("SELECT a FROM nowhere", P.SQLException, "No table named 'nowhere'"),
    ("SELECT nope FROM employee", P.SQLException, "No column named 'nope'"),
    ("SELECT state FROM employee GROUP BY state", P.SQLException, "GROUP BY without aggregate"),
    ("SELECT CAST(salary AS decimal) FROM employee", P.SQLException, "Invalid data type decimal"),
    ("SELECT AVG(salary) FROM employee", P.SQLException, "Invalid aggregate function"),
])
def test_planner_errors_are_the_references(ctx, sql, exc, match):
    with pytest.raises(exc, match=match):
        df = ctx.sql(sql)
        df.schema()
        list(ctx.execute(df))


def test_number_format_error_surfaces_from_the_cast(ctx, oracle):
    with pytest.raises(oracle.OracleError, match="NumberFormat"):
        list(ctx.execute(ctx.sql("SELECT MAX(CAST(first_name AS double)) FROM employee")))


# ---------------------------------------------------------------- the extensions, against direct operator calls
SPECS2 = [dict(kind=2, col_id=i, flo=0.0, fhi=1.0) for i in range(3)]
SPECS3 = [dict(kind=5, col_id=0, dict="ALAKAZARCACOCTDE", dict_width=2), dict(kind=2, col_id=1, flo=0.0, fhi=1000.0)]


def memory_table(oracle, specs, names, types, n, batches):
    step = n // batches
    data = [oracle.generate(specs, 42, i * step, (i + 1) * step) for i in range(batches)]
    return P.InMemoryDataSource(oracle, P.Schema([P.Field(nm, t) for nm, t in zip(names, types)]), data), oracle.generate(specs, 42, 0, n)


def test_baseline_config2_query_string(oracle):
    src, whole = memory_table(oracle, SPECS2, ["a", "b", "c"], [P.DoubleType] * 3, 60_000, 3)
    c = P.ExecutionContext(oracle)
    c.registerDataSource("t", src)
    df = c.sql("SELECT a * b + c AS r FROM t WHERE a > 0.5 AND b < 0.5")
    assert df.schema() == P.Schema([P.Field("r", P.DoubleType)])
    got = [v for b in c.execute(df) for v in b.to_arrow()[0].to_pylist()]
    E = oracle
    pred = E.binary("AND", E.binary("GT", E.col(0), E.lit_f64(0.5)), E.binary("LT", E.col(1), E.lit_f64(0.5)))
    want = E.filter_project(pred, [E.binary("ADD", E.binary("MUL", E.col(0), E.col(1)), E.col(2))], whole).to_arrow()[0].to_pylist()
    assert got == want and 0 < len(got) < 60_000


def test_baseline_config3_query_string(oracle):
    src, whole = memory_table(oracle, SPECS3, ["state", "v"], [P.StringType, P.DoubleType], 40_000, 4)
    c = P.ExecutionContext(oracle)
    c.registerDataSource("t", src)
    df = c.sql("SELECT state, SUM(v), MIN(v), MAX(v), COUNT(v) AS n FROM t GROUP BY state")
    assert df.schema().names() == ["state", "SUM", "MIN", "MAX", "n"] and df.schema().fields[4].dataType == P.Int64Type
    got = sorted(rows(c.execute(df)))
    E = oracle
    agg = E.HashAggregate([E.col(0)], [("SUM", E.col(1)), ("MIN", E.col(1)), ("MAX", E.col(1)), ("COUNT", E.col(1))])
    agg.update(whole)
    want = sorted(rows([agg.finalize()]))
    assert len(got) == 8 and [r[0] for r in got] == [r[0] for r in want]
    for g, w in zip(got, want):
        assert g[2:] == w[2:] and abs(g[1] - w[1]) <= 1e-9 * abs(w[1])            # SUM order differs with the batch boundaries


def test_where_below_an_aggregate_is_fused_and_order_of_select_list_is_kept(oracle):
    src, whole = memory_table(oracle, SPECS3, ["state", "v"], [P.StringType, P.DoubleType], 20_000, 2)
    c = P.ExecutionContext(oracle)
    c.registerDataSource("t", src)
    df = c.sql("SELECT COUNT(v) AS n, state, MAX(v) FROM t WHERE v >= 500.0 GROUP BY state")
    assert df.schema().names() == ["n", "state", "MAX"]
    phys = P.createPhysicalPlan(P.ProjectionPushDownRule().optimize(df.logicalPlan()), oracle)
    agg = phys.children()[0]
    assert type(agg).__name__ == "HashAggregateExec" and agg.predicate is not None and type(agg.children()[0]).__name__ == "ScanExec"
    got = {r[1]: (r[0], r[2]) for r in rows(c.execute(df))}
    E = oracle
    a = E.HashAggregate([E.col(0)], [("COUNT", E.col(1)), ("MAX", E.col(1))], pred=E.binary("GE", E.col(1), E.lit_f64(500.0)))
    a.update(whole)
    assert got == {r[0]: (r[1], r[2]) for r in rows([a.finalize()])}


def test_print_query_result(ctx, capsys):
    P.printQueryResult(ctx.execute(ctx.sql("SELECT id, last_name FROM employee")))
    assert capsys.readouterr().out == "1 Johansson \n2 Person \n3 Pärsson \n"


def test_zero_input_rows_give_one_batch_without_rows(oracle):
    """Rule R10: HashAggregateExec always yields exactly one batch, also for no input (and for a global aggregate)."""
    c = P.ExecutionContext(oracle)
    c.registerDataSource("t", P.InMemoryDataSource(oracle, P.Schema([P.Field("v", P.DoubleType)]), []))
    out = list(c.execute(c.sql("SELECT MAX(v) FROM t")))
    assert len(out) == 1 and out[0].row_count() == 0


# ---------------------------------------------------------------- generated queries against numpy
import numpy as np
from hypothesis import given, settings, strategies as st

_LEAF = st.one_of(st.sampled_from(["a", "b", "c"]), st.floats(0.0, 2.0, allow_nan=False).map(lambda v: round(v, 3)))
_ARITH = st.recursive(_LEAF, lambda kids: st.tuples(st.sampled_from(["+", "-", "*", "/"]), kids, kids), max_leaves=5)
_CMP = st.tuples(st.sampled_from(["<", "<=", ">", ">=", "=", "!="]), _ARITH, _ARITH)
_PRED = st.recursive(_CMP, lambda kids: st.tuples(st.sampled_from(["AND", "OR"]), kids, kids), max_leaves=4)
_PREC = {"OR": 20, "AND": 30, "<": 40, "<=": 40, ">": 40, ">=": 40, "=": 40, "!=": 40, "+": 50, "-": 50, "*": 60, "/": 60}


def _sql(t, parent=0, right=False):
    """The tree as SQL with only the parentheses precedence and left-associativity require."""
    if isinstance(t, str):
        return t
    if isinstance(t, float):
        return repr(t)
    op, l, r = t
    p = _PREC[op]
    s = f"{_sql(l, p)} {op} {_sql(r, p, True)}"
    return f"({s})" if p < parent or (p == parent and right) else s


def _np(t, cols):
    if isinstance(t, str):
        return cols[t]
    if isinstance(t, float):
        return np.full_like(cols["a"], t)
    op, l, r = t
    x, y = _np(l, cols), _np(r, cols)
    with np.errstate(all="ignore"):
        return {"+": np.add, "-": np.subtract, "*": np.multiply, "/": np.divide, "<": np.less, "<=": np.less_equal, ">": np.greater,
                ">=": np.greater_equal, "=": np.equal, "!=": np.not_equal, "AND": np.logical_and, "OR": np.logical_or}[op](x, y)


@settings(max_examples=120, deadline=None)
@given(proj=_ARITH.filter(lambda t: not isinstance(t, float)), pred=_PRED)
def test_generated_select_where_matches_numpy(oracle, proj, pred):
    """SELECT <arithmetic> FROM t WHERE <predicate>: parser precedences and associativity, the planner, and the expression
    rules E2-E4 (separately rounded IEEE arithmetic, comparisons, AND/OR) against numpy on the same columns, bit for bit."""
    batch = oracle.generate(SPECS2, 7, 0, 2000)
    cols = dict(zip("abc", [a.to_numpy(zero_copy_only=False) for a in batch.to_arrow()]))
    c = P.ExecutionContext(oracle)
    c.registerDataSource("t", P.InMemoryDataSource(oracle, P.Schema([P.Field(n, P.DoubleType) for n in "abc"]), [batch]))
    sql = f"SELECT {_sql(proj)} AS r FROM t WHERE {_sql(pred)}"
    got = np.array([v for b in c.execute(c.sql(sql)) for v in b.to_arrow()[0].to_pylist()], dtype=np.float64)
    want = _np(proj, cols)[_np(pred, cols)]
    assert got.tobytes() == want.astype(np.float64).tobytes(), sql


# ---------------------------------------------------------------- from the query string to the sm_100a kernel, without a device
class _Described(P.DataSource):
    """A table that is only a schema: enough to plan and to generate kernels."""

    def __init__(self, fields):
        self._schema = P.Schema([P.Field(n, t) for n, t in fields])

    def schema(self):
        return self._schema

    def scan(self, projection):
        raise AssertionError("planning must not scan")


@pytest.fixture(scope="module")
def X():
    import build
    build.build()
    import kqgpu
    return kqgpu.Exprs()          # the product's expression factory and code generator: host objects, no device needed


def test_baseline_query_strings_compile_to_kernels(X):
    c = P.ExecutionContext(X)
    c.registerDataSource("employee", _Described([(n, P.StringType) for n in ("id", "first_name", "last_name", "state", "job_title", "salary")]))
    c.registerDataSource("t2", _Described([(n, P.DoubleType) for n in "abc"]))
    c.registerDataSource("t3", _Described([("state", P.StringType), ("v", P.DoubleType)]))
    c.registerDataSource("lineitem", _Described([("l_shipdate", P.Date32Type), ("l_returnflag", P.StringType), ("l_linestatus", P.StringType)] +
                                                [(n, P.DoubleType) for n in ("l_quantity", "l_extendedprice", "l_discount", "l_tax")]))

    def kernels(sql):
        phys = P.createPhysicalPlan(P.ProjectionPushDownRule().optimize(c.sql(sql).logicalPlan()), X)
        return phys, P.explain(phys)

    # configs[0]: one fused filter+projection kernel with the Utf8 comparison, five Utf8 columns passed through
    phys, ks = kernels("SELECT id, first_name, last_name, state, salary FROM employee WHERE state = 'CO'")
    assert [type(n).__name__ for n, _ in ks] == ["ProjectionExec"] and "KQ_KERNEL_FILTER" in ks[0][1]
    # configs[1]: a*b+c as two separately rounded operations, never an FMA
    _, ks = kernels("SELECT a * b + c FROM t2 WHERE a > 0.5 AND b < 0.5")
    assert "__dmul_rn" in ks[0][1] and "__dadd_rn" in ks[0][1] and "fma" not in ks[0][1].lower()
    # configs[2]: the aggregate kernel, then the projection that orders the select list (bare columns: aliases, no kernel work)
    _, ks = kernels("SELECT state, SUM(v), MIN(v), MAX(v), COUNT(v) FROM t3 GROUP BY state")
    assert [type(n).__name__ for n, _ in ks] == ["HashAggregateExec"]
    # the reference's own query (Main.kt:1336): the cast is fused into the aggregate kernel
    _, ks = kernels("SELECT state, MAX(CAST(salary AS double)) AS max_amount FROM employee GROUP BY state")
    assert type(ks[0][0]).__name__ == "HashAggregateExec" and len(ks[0][1]) > 1000
    # configs[4], TPC-H Q1 shape: date filter, derived projections, two group keys — by SQL and through the DataFrame API
    q1_sql = ("SELECT l_returnflag, l_linestatus, SUM(l_quantity), SUM(l_extendedprice), SUM(l_extendedprice * (1.0 - l_discount)), "
              "SUM(l_extendedprice * (1.0 - l_discount) * (1.0 + l_tax)), COUNT(l_quantity) FROM lineitem "
              "WHERE l_shipdate <= DATE '1998-09-02' GROUP BY l_returnflag, l_linestatus")
    phys_sql, ks = kernels(q1_sql)
    li = c.sql("SELECT l_returnflag FROM lineitem").logicalPlan().children()[0]
    one = P.lit(1.0)
    disc_price = P.BinaryExpr("MUL", P.col("l_extendedprice"), P.BinaryExpr("SUB", one, P.col("l_discount")))
    charge = P.BinaryExpr("MUL", disc_price, P.BinaryExpr("ADD", one, P.col("l_tax")))
    q1 = P.DataFrame(li).filter(P.BinaryExpr("LE", P.col("l_shipdate"), P.Literal(10471, P.Date32Type))).aggregate(
        [P.col("l_returnflag"), P.col("l_linestatus")],
        [P.Sum(P.col("l_quantity")), P.Sum(P.col("l_extendedprice")), P.Sum(disc_price), P.Sum(charge), P.Count(P.col("l_quantity"))])
    phys = P.createPhysicalPlan(P.ProjectionPushDownRule().optimize(q1.logicalPlan()), X)
    assert type(phys).__name__ == "HashAggregateExec" and phys.predicate is not None            # the date filter runs inside the aggregate kernel
    assert str(phys_sql.children()[0].children()[0]) == str(phys.children()[0])                   # the same scan below both
    assert [src for _, src in ks] == [src for _, src in P.explain(phys)]                           # and the same kernel: 1998-09-02 = day 10471


def test_baseline_config5_query_string_runs_like_the_bench_workload(oracle):
    """configs[4] (TPC-H Q1 shape) by SQL on bench.py's synthetic lineitem, against the operator calls bench.py itself makes."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    wl = bench.WORKLOADS["cfg5"](30_000)
    names = ["l_shipdate", "l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax"]
    types = [P.Date32Type, P.StringType, P.StringType] + [P.DoubleType] * 4
    parts = [oracle.generate(wl.specs(), 42, i * 10_000, (i + 1) * 10_000) for i in range(3)]
    c = P.ExecutionContext(oracle)
    c.registerDataSource("lineitem", P.InMemoryDataSource(oracle, P.Schema([P.Field(n, t) for n, t in zip(names, types)]), parts))
    df = c.sql("SELECT l_returnflag, l_linestatus, SUM(l_quantity), SUM(l_extendedprice), SUM(l_extendedprice * (1.0 - l_discount)), "
               "SUM(l_extendedprice * (1.0 - l_discount) * (1.0 + l_tax)), COUNT(l_quantity) FROM lineitem "
               "WHERE l_shipdate <= DATE '1998-09-02' GROUP BY l_returnflag, l_linestatus")
    got = sorted(rows(c.execute(df)))
    E = oracle
    one = E.lit_f64(1.0)
    dp = E.binary("MUL", E.col(4), E.binary("SUB", one, E.col(5)))
    ch = E.binary("MUL", dp, E.binary("ADD", one, E.col(6)))
    agg = E.HashAggregate([E.col(1), E.col(2)], [("SUM", E.col(3)), ("SUM", E.col(4)), ("SUM", dp), ("SUM", ch), ("COUNT", E.col(3))],
                          pred=E.binary("LE", E.col(0), E.lit_date32(10471)))
    agg.update(oracle.generate(wl.specs(), 42, 0, 30_000))
    want = sorted(rows([agg.finalize()]))
    assert len(got) == len(want) == 6
    for g, w in zip(got, want):
        assert g[:2] == w[:2] and g[6] == w[6] and g[2] == w[2]              # keys, COUNT, and the integer-valued SUM(l_quantity): exact
        assert all(abs(a - b) <= 1e-9 * abs(b) for a, b in zip(g[3:6], w[3:6]))


_ILEAF = st.one_of(st.sampled_from(["a", "b", "c"]), st.integers(0, 1 << 40))
_IARITH = st.recursive(_ILEAF, lambda kids: st.tuples(st.sampled_from(["+", "-", "*"]), kids, kids), max_leaves=5)
_ICMP = st.tuples(st.sampled_from(["<", "<=", ">", ">=", "=", "!="]), _IARITH, _IARITH)


def _isql(t, parent=0, right=False):
    if isinstance(t, (str, int)):
        return str(t)
    op, l, r = t
    p = _PREC[op]
    s = f"{_isql(l, p)} {op} {_isql(r, p, True)}"
    return f"({s})" if p < parent or (p == parent and right) else s


def _inp(t, cols):
    if isinstance(t, str):
        return cols[t]
    if isinstance(t, int):
        return np.full_like(cols["a"], t)
    op, l, r = t
    x, y = _inp(l, cols), _inp(r, cols)
    with np.errstate(all="ignore"):
        return {"+": np.add, "-": np.subtract, "*": np.multiply, "<": np.less, "<=": np.less_equal, ">": np.greater,
                ">=": np.greater_equal, "=": np.equal, "!=": np.not_equal}[op](x, y)


@settings(max_examples=80, deadline=None)
@given(proj=_IARITH.filter(lambda t: not isinstance(t, int)), pred=_ICMP)
def test_generated_int64_queries_wrap_like_kotlin_long(oracle, proj, pred):
    """Rule E1: Int64 arithmetic wraps in two's complement like Kotlin's Long (numpy's int64 does the same), through the
    SQL front-end; products of 40-bit operands overflow often."""
    specs = [dict(kind=1, col_id=i, ilo=0, ihi=1 << 40) for i in range(3)]
    batch = oracle.generate(specs, 9, 0, 1500)
    cols = dict(zip("abc", [a.to_numpy(zero_copy_only=False).astype(np.int64) for a in batch.to_arrow()]))
    c = P.ExecutionContext(oracle)
    c.registerDataSource("t", P.InMemoryDataSource(oracle, P.Schema([P.Field(n, P.Int64Type) for n in "abc"]), [batch]))
    sql = f"SELECT {_isql(proj)} AS r FROM t WHERE {_isql(pred)}"
    got = np.array([v for b in c.execute(c.sql(sql)) for v in b.to_arrow()[0].to_pylist()], dtype=np.int64)
    want = _inp(proj, cols)[_inp(pred, cols)]
    assert np.array_equal(got, want), sql


def test_int64_division_truncates_and_by_zero_is_arithmetic_exception(oracle):
    """Rule E4: Kotlin's Long division truncates toward zero; dividing by zero throws ArithmeticException — only for rows
    that pass the filter, as in a row-at-a-time engine."""
    import pyarrow as pa
    batch = oracle.RecordBatch.from_arrow([pa.array([7, -7, 7, -7, 5], pa.int64()), pa.array([2, 2, -2, -2, 0], pa.int64())])
    c = P.ExecutionContext(oracle)
    c.registerDataSource("t", P.InMemoryDataSource(oracle, P.Schema([P.Field("x", P.Int64Type), P.Field("y", P.Int64Type)]), [batch]))
    assert rows(c.execute(c.sql("SELECT x / y FROM t WHERE y != 0"))) == [(3,), (-3,), (-3,), (3,)]
    with pytest.raises(oracle.OracleError, match="Arithmetic"):
        list(c.execute(c.sql("SELECT x / y FROM t")))


def test_in_memory_source_checks_its_schema(oracle):
    batch = oracle.generate([dict(kind=2, col_id=0, flo=0.0, fhi=1.0)], 1, 0, 10)
    with pytest.raises(P.IllegalStateException, match="the schema says Int\\(64, true\\)"):
        P.InMemoryDataSource(oracle, P.Schema([P.Field("v", P.Int64Type)]), [batch])
    with pytest.raises(P.IllegalStateException, match="columns under a schema"):
        P.InMemoryDataSource(oracle, P.Schema([P.Field("v", P.DoubleType), P.Field("w", P.DoubleType)]), [batch])
