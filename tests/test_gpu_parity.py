"""GPU parity tests: libkqgpu.so (through the C ABI) against the CPU oracle on the same inputs.

Bit-exact for integer, boolean, string, row-count and group-key outputs and for Float64 results of
expressions (no FMA contraction); Float64 SUMs within 1e-9 relative (reassociated reduction order,
north_star); exact when the addends are integer-valued.
"""
import json
import math
import os

import numpy as np
import pyarrow as pa
import pytest

from planspec import b, build, col, lit, sort_rows

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "employee_golden.json")
F64, UTF8, I64, BOOL, DATE32 = 1, 2, 3, 4, 5
SUM_RTOL = 1e-9     # north_star: float64 sums within 1e-9 relative tolerance


@pytest.fixture(scope="module")
def G(gpu, gctx):
    return gpu.Engine(gctx)


def same(a: pa.Array, c: pa.Array):
    """bit-exact equality of two arrow arrays, NaN == NaN, -0.0 != +0.0."""
    assert a.type == c.type, (a.type, c.type)
    assert len(a) == len(c), (len(a), len(c))
    la, lc = a.to_pylist(), c.to_pylist()
    if a.type == pa.float64():
        ka = [None if v is None else np.float64(v).tobytes() if v == v else "nan" for v in la]
        kc = [None if v is None else np.float64(v).tobytes() if v == v else "nan" for v in lc]
        assert ka == kc
    else:
        assert la == lc


def rand_table(rng, n, null_frac):
    def mask(x, t):
        return pa.array(x, type=t, mask=rng.random(n) < null_frac) if null_frac else pa.array(x, type=t)
    return [mask(rng.integers(0, 1 << 20, n), pa.int64()), mask(rng.integers(0, 1 << 20, n), pa.int64()),
            mask(rng.integers(-50, 50, n), pa.int64()),
            mask(rng.random(n), pa.float64()), mask(rng.random(n), pa.float64()), mask(np.floor(rng.random(n) * 1000), pa.float64()),
            mask(np.array(["AL", "AK", "AZ", "CO", "NY", "Uppsala", ""], dtype=object)[rng.integers(0, 7, n)], pa.string()),
            mask(rng.random(n) < 0.5, pa.bool_()),
            mask(rng.integers(8000, 11000, n).astype(np.int32), pa.int32()).cast(pa.date32())]


SPECS = [
    b("ADD", b("MUL", col(0), col(1)), col(2)),
    b("ADD", b("MUL", col(3), col(4)), col(5)),
    b("AND", b("GT", col(0), lit("i64", 1 << 19)), b("LT", col(1), lit("i64", 1 << 19))),
    b("AND", b("GT", col(3), lit("f64", 0.5)), b("LT", col(4), lit("f64", 0.5))),
    b("OR", b("EQ", col(6), lit("utf8", "CO")), b("AND", col(7), b("GE", col(5), lit("f64", 500.0)))),
    b("SUB", col(2), b("DIV", col(0), b("ADD", col(2), lit("i64", 100)))),
    b("DIV", col(3), b("SUB", col(4), lit("f64", 0.25))),
    b("NE", col(6), col(6)),
    b("LE", col(2), lit("i64", None)),
    b("LE", col(8), lit("date32", 9500)),
    b("LT", lit("utf8", "AZ"), col(6)),
    b("GE", col(6), lit("utf8", "CO")),
    ("cast", col(2), "f64"),
    b("MUL", b("SUB", lit("f64", 1.0), col(3)), b("ADD", lit("f64", 1.0), col(4))),
]


# ---------------------------------------------------------------- golden vectors (employee.csv)
@pytest.fixture(scope="module")
def employee():
    with open(GOLDEN, encoding="utf-8") as f:
        return json.load(f)


def employee_batch(E, g):
    return E.RecordBatch.from_arrow([pa.array(g["columns"][name], pa.string()) for name in g["schema"]])


def test_golden_roundtrip_and_group_by(G, employee):
    batch = employee_batch(G, employee)
    assert batch.row_count() == 3 and batch.num_columns() == 6
    for name, arr in zip(employee["schema"], batch.to_arrow()):
        assert arr.to_pylist() == employee["columns"][name]
    s = employee["schema"]
    agg = G.HashAggregate([G.col(s.index("state"))], [("MAX", G.cast(G.col(s.index("salary")), F64))])
    agg.update(batch)
    keys, mx = agg.finalize().to_arrow()
    assert dict(zip(keys.to_pylist(), mx.to_pylist())) == employee["group_by_state_max_salary"]


@pytest.mark.parametrize("state,key", [("CO", "config1_where_state_eq_CO"), ("Uppsala", "where_state_eq_Uppsala")])
def test_golden_config1_filter_project(G, employee, state, key):
    batch = employee_batch(G, employee)
    s = employee["schema"]
    names = ["id", "first_name", "last_name", "state", "salary"]
    pred = G.binary("EQ", G.col(s.index("state")), G.lit_utf8(state))
    out = G.filter_project(pred, [G.col(s.index(n)) for n in names], batch)
    got = {n: a.to_pylist() for n, a in zip(names, out.to_arrow())}
    assert got == employee[key] and out.row_count() == len(employee[key]["id"])


# ---------------------------------------------------------------- expressions / projection
@pytest.mark.parametrize("n", [0, 1, 1023, 1024, 1025, 70001])
@pytest.mark.parametrize("null_frac", [0.0, 0.05])
def test_projection_matches_oracle(G, oracle, n, null_frac):
    rng = np.random.default_rng(42 + n)
    arrs = rand_table(rng, n, null_frac)
    gb, ob = G.RecordBatch.from_arrow(arrs), oracle.RecordBatch.from_arrow(arrs)
    for lo in range(0, len(SPECS), 7):
        chunk = SPECS[lo:lo + 7]
        try:
            want = oracle.project([build(oracle, s) for s in chunk], ob).to_arrow()
        except oracle.OracleError as e:
            assert e.code == 6          # the DIV spec may hit / by zero
            with pytest.raises(G.KqError) as ge:
                G.project([build(G, s) for s in chunk], gb).to_arrow()
            assert ge.value.code == 6
            chunk = [s for s in chunk if s is not SPECS[5]]
            want = oracle.project([build(oracle, s) for s in chunk], ob).to_arrow()
        got = G.project([build(G, s) for s in chunk], gb)
        assert got.row_count() == n
        for a, w in zip(got.to_arrow(), want):
            same(a, w)


def test_column_expression_is_alias(G):
    batch = G.RecordBatch.from_arrow([pa.array([1.0, 2.0]), pa.array(["x", None])])
    c = G.col(1).evaluate(batch)
    assert c.to_arrow().to_pylist() == ["x", None]
    assert c.device_ptrs() == batch.field(1).device_ptrs()      # zero copy (Main.kt:453-455)
    assert c.get_value(0) == "x" and c.get_value(1) is None     # ColumnVector.getValue


def test_no_fma_contraction(G):
    a, bb, c = 1.0 + 2.0 ** -30, 1.0 + 2.0 ** -30, -(1.0 + 2.0 ** -29)
    batch = G.RecordBatch.from_arrow([pa.array([a] * 5), pa.array([bb] * 5), pa.array([c] * 5)])
    got = G.binary("ADD", G.binary("MUL", G.col(0), G.col(1)), G.col(2)).evaluate(batch).to_arrow().to_pylist()
    assert got == [(a * bb) + c] * 5 and got[0] == 0.0          # fma would give 2^-60


def test_error_classes_match_the_oracle(G, oracle):
    arrs = [pa.array([1, 2]), pa.array([1.0, 2.0]), pa.array(["a", "b"])]
    cases = [
        (lambda E: E.binary("ADD", E.col(0), E.col(1)), 1),     # operand types differ
        (lambda E: E.binary("AND", E.col(0), E.col(0)), 1),     # AND on non-Bool
        (lambda E: E.col(7), 1),                                # field index out of range
        (lambda E: E.cast(E.col(0), UTF8), 1),                  # cast target (Main.kt:799)
        (lambda E: E.cast(E.binary("EQ", E.col(0), E.col(0)), F64), 1),   # Cannot cast value (Main.kt:792)
        (lambda E: E.binary("DIV", E.col(0), E.lit_i64(0)), 6), # ArithmeticException
        (lambda E: E.binary("MUL", E.col(2), E.col(2)), None),  # math on Utf8: both must fail
    ]
    for mk, code in cases:
        with pytest.raises(oracle.OracleError) as oe:
            mk(oracle).evaluate(oracle.RecordBatch.from_arrow(arrs)).to_arrow()
        with pytest.raises(G.KqError) as ge:
            mk(G).evaluate(G.RecordBatch.from_arrow(arrs)).to_arrow()
        if code is not None:
            assert oe.value.code == code and ge.value.code == code
    G.ctx.sync()   # the context stays usable after errors
    assert G.col(0).evaluate(G.RecordBatch.from_arrow(arrs)).to_arrow().to_pylist() == [1, 2]


def test_cast_utf8_to_double(G, oracle):
    vals = ["1337", " 6.5e1 ", None, "-0", "NaN", "Infinity", "-Infinity", "1d", ".5", "1.", "0.1", "52.5", "-3.00",
            "123456789012345", "1e22", "1e-22", "9007199254740992", "0.000001", "+7"]
    arr = [pa.array(vals)]
    got = G.cast(G.col(0), F64).evaluate(G.RecordBatch.from_arrow(arr)).to_arrow()
    want = oracle.cast(oracle.col(0), F64).evaluate(oracle.RecordBatch.from_arrow(arr)).to_arrow()
    same(got, want)
    for bad in ["", "abc", "1_0", "nan", "inf", "1e", "--1", "1 2", "0x1", "0x.p1"]:
        with pytest.raises(G.KqError) as e:
            G.cast(G.col(0), F64).evaluate(G.RecordBatch.from_arrow([pa.array([bad])])).to_arrow()
        assert e.value.code == 5, bad


def test_cast_utf8_to_double_is_exact_for_every_valid_input(G, oracle):
    """String.toDouble() has no digit or exponent limit (Main.kt:791): inputs outside the one-multiply fast path take the
    exact big-integer tier (csrc/kq_parse.cuh) on the device. 200 k random decimals (1-40 digits, a few up to 900,
    exponents +-340), halfway cases, subnormals, overflow, hex floats — bit for bit against the oracle."""
    import random
    rng = random.Random(11)
    vals = ["1e23", "9007199254740993", "123456789012345678901234567890", "0.1e-5000", "1e5000", "4.9e-324", "2.4703282292062327e-324",
            "2.4703282292062328e-324", "1.7976931348623159e308", "1.7976931348623157e308", "0x1.8p1", "0x1p-1075", "0x1.0000000000001p-1075",
            "0x1.fffffffffffff8p1023", "0xAbC.dEfp-7f", "-0x1p3D", "8.5e-320", "1" + "0" * 400, "0." + "0" * 400 + "1", "3." + "3" * 800]
    for _ in range(200_000):
        nd = rng.randint(1, 40) if rng.random() < 0.97 else rng.randint(40, 900)
        digs = "".join(rng.choice("0123456789") for _ in range(nd))
        if rng.random() < 0.6:
            k = rng.randint(0, nd)
            digs = digs[:k] + "." + digs[k:]
        s = ("-" if rng.random() < 0.3 else "") + digs
        if rng.random() < 0.7:
            s += rng.choice("eE") + rng.choice(["", "+", "-"]) + str(rng.randint(0, 340))
        vals.append(s)
    arr = [pa.array(vals)]
    got = G.cast(G.col(0), F64).evaluate(G.RecordBatch.from_arrow(arr)).to_arrow()
    want = oracle.cast(oracle.col(0), F64).evaluate(oracle.RecordBatch.from_arrow(arr)).to_arrow()
    same(got, want)


# ---------------------------------------------------------------- filter
@pytest.mark.parametrize("n", [0, 1, 1024, 4097, 200003])
@pytest.mark.parametrize("null_frac", [0.0, 0.05])
def test_filter_project_matches_oracle(G, oracle, n, null_frac):
    rng = np.random.default_rng(7 + n)
    arrs = rand_table(rng, n, null_frac)
    for pred, proj in [(SPECS[2], [SPECS[0], col(6), SPECS[1], SPECS[3]]), (SPECS[4], [col(7), col(8), SPECS[9], col(0)])]:
        want = oracle.filter_project(build(oracle, pred), [build(oracle, p) for p in proj], oracle.RecordBatch.from_arrow(arrs))
        got = G.filter_project(build(G, pred), [build(G, p) for p in proj], G.RecordBatch.from_arrow(arrs))
        assert got.row_count() == want.row_count()
        for a, w in zip(got.to_arrow(), want.to_arrow()):
            same(a, w)


@pytest.mark.parametrize("n,null_frac", [(0, 0.0), (1000, 0.0), (70001, 0.1), (9_500_000, 0.0), (8_400_000, 0.02)])
def test_filter_project_host_pipeline(G, oracle, n, null_frac):
    """kq_filter_project_host streams host buffers in 4 Mi-row chunks (H2D / kernel / D2H overlapped); the
    compacted output must be the same rows in the same order as one device-resident call and as the oracle."""
    rng = np.random.default_rng(5 + n)
    def mask(x, t):
        return pa.array(x, type=t, mask=rng.random(n) < null_frac) if null_frac else pa.array(x, type=t)
    arrs = [mask(rng.random(n), pa.float64()), mask(rng.random(n), pa.float64()), mask(np.floor(rng.random(n) * 1000), pa.float64()),
            mask(rng.integers(-1000, 1000, n), pa.int64())]
    pred = b("AND", b("GT", col(0), lit("f64", 0.5)), b("LT", col(1), lit("f64", 0.5)))
    proj = [b("ADD", b("MUL", col(0), col(1)), col(2)), b("MUL", col(3), col(3)), b("GE", col(2), lit("f64", 500.0))]
    want = oracle.filter_project(build(oracle, pred), [build(oracle, p) for p in proj], oracle.RecordBatch.from_arrow(arrs)).to_arrow()
    cols = []
    for a in arrs:
        v, d = a.buffers()
        cols.append((G.type_of(a), v.address if v is not None else None, d.address))
    out_d = [np.zeros(max(n, 1), dtype=np.float64), np.zeros(max(n, 1), dtype=np.int64), np.zeros((max(n, 1) + 63) // 64 * 8, dtype=np.uint8)]
    out_v = [np.zeros((max(n, 1) + 63) // 64 * 8, dtype=np.uint8) for _ in proj]
    m = G.filter_project_host(build(G, pred), [build(G, p) for p in proj], cols, n, [(d.ctypes.data, v.ctypes.data) for d, v in zip(out_d, out_v)])
    assert m == len(want[0])
    def valid(k):
        return np.unpackbits(out_v[k], bitorder="little")[:m].astype(bool)
    got = [pa.array(out_d[0][:m], mask=~valid(0)), pa.array(out_d[1][:m], mask=~valid(1)),
           pa.array(np.unpackbits(out_d[2], bitorder="little")[:m].astype(bool), mask=~valid(2))]
    for a, w in zip(got, want):
        same(a, w)


def test_upload_of_a_sliced_utf8_vector_rebases_offsets(gpu, gctx):
    """Arrow slices keep the parent's data buffer and a window of its offsets (offsets[0] != 0)."""
    import ctypes as C
    n = 300_000
    whole = pa.array([("s%d" % (i % 977)) for i in range(n)], type=pa.string())
    _, off, data = whole.buffers()
    for lo, hi in [(0, n), (1, 100), (123_457, n), (299_999, n)]:
        out = C.c_void_p()
        gctx.check(gpu.lib().kq_column_upload(gctx.h, UTF8, hi - lo, None, off.address + 4 * lo, data.address, data.size, C.byref(out)))
        same(gpu.Column(gctx, out).to_arrow(), pa.concat_arrays([whole.slice(lo, hi - lo)]))


@pytest.mark.parametrize("threshold", [-1.0, 0.37, 2.0])         # keeps every row / some rows / no row
def test_filter_selectivity_extremes_and_wide_projection(G, oracle, threshold):
    """A projection wider than one launch's stash is split into several launches that repeat the predicate;
    the outputs must still line up row by row."""
    rng = np.random.default_rng(3)
    n = 150_001
    arrs = [pa.array(rng.random(n)), pa.array(rng.integers(-9, 9, n), mask=rng.random(n) < 0.1), pa.array(rng.random(n) < 0.5)]
    pred = b("GT", col(0), lit("f64", threshold))
    proj = [b("ADD", col(0), lit("f64", float(i))) for i in range(7)] + [b("MUL", col(1), lit("i64", i + 2)) for i in range(5)] + [col(2), b("LT", col(0), lit("f64", 0.5))]
    want = oracle.filter_project(build(oracle, pred), [build(oracle, p) for p in proj], oracle.RecordBatch.from_arrow(arrs))
    got = G.filter_project(build(G, pred), [build(G, p) for p in proj], G.RecordBatch.from_arrow(arrs))
    assert got.row_count() == want.row_count() == (n if threshold < 0 else (0 if threshold > 1 else want.row_count()))
    for a, w in zip(got.to_arrow(), want.to_arrow()):
        same(a, w)


def test_filter_gathers_every_column_in_order(G, oracle):
    rng = np.random.default_rng(11)
    arrs = rand_table(rng, 50001, 0.1)
    for pred in (SPECS[2], SPECS[4], b("EQ", col(6), lit("utf8", "nope")), b("GE", col(0), lit("i64", 0))):
        want, wsel = oracle.filter(build(oracle, pred), oracle.RecordBatch.from_arrow(arrs), want_selection=True)
        got, gsel = G.filter(build(G, pred), G.RecordBatch.from_arrow(arrs), want_selection=True)
        same(gsel.to_arrow(), wsel.to_arrow())
        assert got.num_columns() == len(arrs)
        for a, w in zip(got.to_arrow(), want.to_arrow()):
            same(a, w)


def test_filter_then_downstream_operators(G, oracle):
    """A filtered batch (row count still device-resident) feeds projection and aggregation."""
    rng = np.random.default_rng(5)
    arrs = rand_table(rng, 30000, 0.05)
    def run(E):
        f = E.filter(build(E, SPECS[2]), E.RecordBatch.from_arrow(arrs))
        p = E.project([build(E, SPECS[0]), E.col(6)], f)
        agg = E.HashAggregate([E.col(1)], [("SUM", E.col(0)), ("COUNT", E.col(0))])
        agg.update(p)
        return p.to_arrow(), sort_rows(agg.finalize().to_arrow(), 1)
    (gp, ga), (op, oa) = run(G), run(oracle)
    for a, w in zip(gp, op):
        same(a, w)
    assert ga == oa


def test_div_by_zero_only_counts_on_rows_that_pass_the_filter(G, oracle):
    a = pa.array([10, 20, 30, 40]); d = pa.array([2, 0, 5, 0])
    pred = b("NE", col(1), lit("i64", 0))
    for E in (G, oracle):
        out = E.filter_project(build(E, pred), [build(E, b("DIV", col(0), col(1)))], E.RecordBatch.from_arrow([a, d]))
        assert out.to_arrow()[0].to_pylist() == [5, 6]
    with pytest.raises(G.KqError) as e:
        G.filter_project(build(G, b("GE", col(1), lit("i64", 0))), [build(G, b("DIV", col(0), col(1)))],
                         G.RecordBatch.from_arrow([a, d])).to_arrow()
    assert e.value.code == 6


# ---------------------------------------------------------------- hash aggregate
AGGS = [("SUM", 5), ("MIN", 5), ("MAX", 5), ("COUNT", 5), ("SUM", 2), ("MAX", 0), ("MIN", 8), ("COUNT", 6)]


@pytest.mark.parametrize("n", [0, 1, 1025, 60007])
@pytest.mark.parametrize("null_frac", [0.0, 0.1])
@pytest.mark.parametrize("keys", [[6], [2], [6, 7], [], [8, 2], [3]])
def test_hash_aggregate_matches_oracle(G, oracle, keys, null_frac, n):
    rng = np.random.default_rng(3 + n)
    arrs = rand_table(rng, n, null_frac)
    def run(E):
        agg = E.HashAggregate([E.col(k) for k in keys], [(kind, E.col(c)) for kind, c in AGGS])
        agg.update(E.RecordBatch.from_arrow(arrs))
        return agg.finalize()
    got, want = run(G), run(oracle)
    assert got.row_count() == want.row_count()
    ga, wa = got.to_arrow(), want.to_arrow()
    assert [a.type for a in ga] == [a.type for a in wa]
    assert sort_rows(ga, len(keys)) == sort_rows(wa, len(keys))


def test_hash_aggregate_float_sums_within_tolerance(G, oracle):
    rng = np.random.default_rng(9)
    n = 300000
    arrs = [pa.array(rng.integers(0, 50, n)), pa.array(rng.random(n) * 1000)]
    def run(E):
        agg = E.HashAggregate([E.col(0)], [("SUM", E.col(1)), ("MIN", E.col(1)), ("MAX", E.col(1)), ("COUNT", E.col(1))])
        agg.update(E.RecordBatch.from_arrow(arrs))
        return sort_rows(agg.finalize().to_arrow(), 1)
    got, want = run(G), run(oracle)
    assert len(got) == len(want) == 50
    for g, w in zip(got, want):
        assert g[0] == w[0] and g[2:] == w[2:]                       # key, MIN, MAX, COUNT bit-exact
        assert abs(g[1] - w[1]) <= SUM_RTOL * abs(w[1])               # SUM: reassociated


def test_hash_aggregate_is_independent_of_batch_boundaries(G, oracle):
    rng = np.random.default_rng(21)
    arrs = rand_table(rng, 20000, 0.05)
    def run(E, cuts):
        agg = E.HashAggregate([E.col(6), E.col(7)], [(k, E.col(c)) for k, c in AGGS])
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            agg.update(E.RecordBatch.from_arrow([a.slice(lo, hi - lo) for a in arrs]))
        return sort_rows(agg.finalize().to_arrow(), 2)
    assert run(G, [0, 20000]) == run(G, [0, 1, 5000, 5000, 19999, 20000]) == run(oracle, [0, 20000])


def test_high_cardinality_aggregate_grows_the_table(G, oracle):
    rng = np.random.default_rng(33)
    n = 400000
    arrs = [pa.array(rng.integers(0, 150000, n)), pa.array(np.floor(rng.random(n) * 100))]
    def run(E):
        agg = E.HashAggregate([E.col(0)], [("SUM", E.col(1)), ("MIN", E.col(1)), ("MAX", E.col(1)), ("COUNT", E.col(1))])
        agg.update(E.RecordBatch.from_arrow(arrs))
        agg.update(E.RecordBatch.from_arrow(arrs))
        return sort_rows(agg.finalize().to_arrow(), 1)
    assert run(G) == run(oracle)


@pytest.mark.parametrize("null_frac", [0.0, 0.05])
@pytest.mark.parametrize("keys", [[0], [0, 7], [3]])
def test_partitioned_path_matches_oracle(G, oracle, keys, null_frac):
    """High-cardinality hint + a batch above the row threshold: the partitioned path (scatter into hash
    partitions, reduce each in shared memory). Int64 / Int64+Bool / Float64 keys, nullable keys and inputs."""
    rng = np.random.default_rng(77)
    n = 1_100_000
    arrs = rand_table(rng, n, null_frac)
    def run(E, **kw):
        agg = E.HashAggregate([E.col(k) for k in keys], [(kind, E.col(c)) for kind, c in AGGS], **kw)
        agg.update(E.RecordBatch.from_arrow(arrs))
        agg.update(E.RecordBatch.from_arrow([a.slice(7, 1_050_000) for a in arrs]))     # a second batch merges into the same table
        return agg.finalize()
    got, want = run(G, expected_groups=700_000), run(oracle)
    assert got.row_count() == want.row_count()
    assert sort_rows(got.to_arrow(), len(keys)) == sort_rows(want.to_arrow(), len(keys))


@pytest.mark.parametrize("null_frac", [0.0, 0.1])
@pytest.mark.parametrize("ngroups,hint", [(700, 700), (5000, 900)])     # fits the block-shared table / overflows it (rows spill to the global table)
def test_mid_cardinality_shared_table_matches_oracle(G, oracle, ngroups, hint, null_frac):
    rng = np.random.default_rng(5)
    n = 150_000
    arrs = rand_table(rng, n, null_frac)
    k = rng.integers(0, ngroups, n)
    arrs[0] = pa.array(k, type=pa.int64(), mask=rng.random(n) < null_frac) if null_frac else pa.array(k, type=pa.int64())
    for keys in ([0], [0, 7]):
        def run(E, **kw):
            agg = E.HashAggregate([E.col(c) for c in keys], [(kind, E.col(c)) for kind, c in AGGS], **kw)
            agg.update(E.RecordBatch.from_arrow(arrs))
            agg.update(E.RecordBatch.from_arrow([a.slice(11, 40_000) for a in arrs]))
            return agg.finalize()
        got, want = run(G, expected_groups=hint), run(oracle)
        assert got.row_count() == want.row_count()
        assert sort_rows(got.to_arrow(), len(keys)) == sort_rows(want.to_arrow(), len(keys))


def test_partitioned_path_survives_skewed_keys(G, oracle):
    """Nine rows in ten share one key: its buckets overflow and the surplus rows take the plain global path."""
    rng = np.random.default_rng(78)
    n = 1_500_000
    k = rng.integers(0, 400_000, n)
    k[rng.random(n) < 0.9] = 123456789
    arrs = [pa.array(k), pa.array(np.floor(rng.random(n) * 1000))]
    def run(E, **kw):
        agg = E.HashAggregate([E.col(0)], [("SUM", E.col(1)), ("MIN", E.col(1)), ("MAX", E.col(1)), ("COUNT", E.col(1))], **kw)
        agg.update(E.RecordBatch.from_arrow(arrs))
        return sort_rows(agg.finalize().to_arrow(), 1)
    assert run(G, expected_groups=400_000) == run(oracle)


def test_fused_filter_project_aggregate_q1_shape(G, oracle):
    """TPC-H Q1 shape (BASELINE config 5) on a small seeded lineitem."""
    specs = q1_specs()
    for E in (G, oracle):
        pass
    def run(E, **kw):
        batch = E.generate(specs, 42, 0, 150000)
        return sort_rows(q1_aggregate(E, batch, **kw).to_arrow(), 2)
    want = run(oracle)
    for kw in ({}, dict(expected_groups=6)):         # the planner's hint picks the tile geometry (bench.py passes 6)
        got = run(G, **kw)
        assert len(got) == len(want) == 6
        for g, w in zip(got, want):
            assert g[:2] == w[:2] and g[6] == w[6]
            for x, y in zip(g[2:6], w[2:6]):
                assert abs(x - y) <= SUM_RTOL * abs(y)


def q1_specs():
    return [dict(kind=6, ilo=8036, ihi=10562),                    # l_shipdate 1992-01-02 .. 1998-12-01
            dict(kind=5, dict="ANR", dict_width=1), dict(kind=5, dict="FO", dict_width=1),
            dict(kind=3, ilo=1, ihi=51),                           # l_quantity
            dict(kind=2, flo=900.0, fhi=105000.0),                 # l_extendedprice
            dict(kind=4, ilo=0, ihi=11, fhi=100.0),                # l_discount
            dict(kind=4, ilo=0, ihi=9, fhi=100.0)]                 # l_tax


def q1_aggregate(E, batch, **kw):
    one = E.lit_f64(1.0)
    disc_price = E.binary("MUL", E.col(4), E.binary("SUB", one, E.col(5)))
    charge = E.binary("MUL", disc_price, E.binary("ADD", one, E.col(6)))
    pred = E.binary("LE", E.col(0), E.lit_date32(10471))          # 1998-09-02
    agg = E.HashAggregate([E.col(1), E.col(2)],
                          [("SUM", E.col(3)), ("SUM", E.col(4)), ("SUM", disc_price), ("SUM", charge), ("COUNT", E.lit_i64(1))],
                          pred=pred, **kw)
    agg.update(batch)
    return agg.finalize()


def test_group_key_equality_and_nan_order(G, oracle):
    """R7 on the GPU: NaN == NaN is one group, +0.0 != -0.0, null is a group. MIN/MAX (rule R9, Main.kt:540-555) are
    checked AGAINST THE ORACLE wherever the reference does not depend on the row order: `value > this.value` never lets a
    NaN replace a held value, an all-NaN group yields NaN. The one order-dependent residual — a NaN as the FIRST non-null
    value sticks in the reference — is asserted separately, with both behaviours spelled out."""
    k = pa.array([float("nan"), 0.0, -0.0, float("nan"), None, None, 1.0], pa.float64())
    v = pa.array([1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0], pa.float64())
    res = {}
    for E in (G, oracle):
        agg = E.HashAggregate([E.col(0)], [("COUNT", E.col(1)), ("MAX", E.col(1))])
        agg.update(E.RecordBatch.from_arrow([k, v]))
        keys, cnt, mx = [a.to_pylist() for a in agg.finalize().to_arrow()]
        got = {}
        for kk, c, m in zip(keys, cnt, mx):
            tag = "null" if kk is None else ("nan" if kk != kk else ("-0" if (kk == 0 and math.copysign(1, kk) < 0) else kk))
            got[tag] = (c, m)
        res[E is G] = got
    assert res[True] == res[False] == {"nan": (2, 4.0), 0.0: (1, 2.0), "-0": (1, 3.0), "null": (2, 6.0), 1.0: (1, 7.0)}

    def mm(E, vals):
        agg = E.HashAggregate([], [("MAX", E.col(0)), ("MIN", E.col(0)), ("SUM", E.col(0)), ("COUNT", E.col(0))])
        agg.update(E.RecordBatch.from_arrow([pa.array(vals, pa.float64())]))
        return [a.to_pylist()[0] for a in agg.finalize().to_arrow()]

    def eqv(a, b):
        return all((x is None and y is None) or (x is not None and y is not None and ((x != x and y != y) or x == y)) for x, y in zip(a, b))

    nan, inf = float("nan"), float("inf")
    # order-independent in the reference: the first non-null value is not a NaN, or every value is
    for vals in ([5.0, nan, -1.0], [-1.0, nan, 5.0, nan], [nan, nan], [nan], [1.0, inf, nan, -inf], [inf, inf], [None, 2.0, nan, None],
                 [None, None], [-0.0], [0.0, 0.0]):
        g, o = mm(G, vals), mm(oracle, vals)
        assert eqv(g[:2], o[:2]) and g[3] == o[3], (vals, g, o)
    assert mm(G, [5.0, nan, -1.0])[:2] == [5.0, -1.0]
    assert mm(G, [None, None]) == [None, None, None, 0]
    # the residual: a leading NaN sticks in the reference (nothing compares greater than it), the GPU ignores every NaN
    o, g = mm(oracle, [nan, 5.0, -1.0]), mm(G, [nan, 5.0, -1.0])
    assert o[0] != o[0] and o[1] != o[1]
    assert g[:2] == [5.0, -1.0]
    # equal-comparing zeros: the reference keeps the first seen (order-dependent); the GPU result compares equal to it
    g, o = mm(G, [0.0, -0.0]), mm(oracle, [0.0, -0.0])
    assert g[0] == o[0] == 0.0 and g[1] == o[1] == 0.0


def test_zero_rows_zero_groups(G):
    agg = G.HashAggregate([], [("MAX", G.col(0)), ("COUNT", G.col(0))])
    agg.update(G.RecordBatch.from_arrow([pa.array([], pa.float64())]))
    out = agg.finalize()
    assert out.row_count() == 0 and out.num_columns() == 2       # rule R10: no SQL "one row" rule


def test_unsupported_aggregates_raise_like_the_reference(G, oracle):
    arrs = [pa.array(["a", "b"]), pa.array([True, False])]
    for c in (0, 1):
        for E in (G, oracle):
            Err = G.KqError if E is G else oracle.OracleError
            with pytest.raises(Err) as e:
                agg = E.HashAggregate([], [("MAX", E.col(c))])
                agg.update(E.RecordBatch.from_arrow(arrs))
                agg.finalize().to_arrow()
            assert e.value.code == 2                                 # UnsupportedOperationException (Main.kt:548)


# ---------------------------------------------------------------- generator
def test_generator_matches_oracle_bit_for_bit(G, oracle):
    specs = [dict(kind=1, ilo=0, ihi=1 << 20), dict(kind=2, flo=0.0, fhi=1000.0, null_per_10k=500),
             dict(kind=3, ilo=0, ihi=1000), dict(kind=4, ilo=0, ihi=11, fhi=100.0),
             dict(kind=5, dict="ALAKAZCO", dict_width=2, null_per_10k=300), dict(kind=6, ilo=8036, ihi=10561),
             dict(kind=7, ilo=2500), dict(kind=2, flo=900.0, fhi=105000.0)]
    for lo, hi in [(0, 0), (0, 1), (5, 70), (1000, 5097), (123456789012, 123456789012 + 3000)]:
        got = G.generate(specs, 42, lo, hi).to_arrow()
        want = oracle.generate(specs, 42, lo, hi).to_arrow()
        for a, w in zip(got, want):
            same(a, w)


# ---------------------------------------------------------------- front-end geometry corner cases
@pytest.mark.parametrize("geom", ["8,7,2", "4,4", "12,6,2", "6,10", "8,3,2,2", "2,2,6,3"])
def test_group_by_is_geometry_independent(G, oracle, geom, monkeypatch):
    """The tile geometry of the aggregate kernel (KQ_AGG_GEOM = "rows per thread, consumer warps[, stages[, CTAs per SM]]",
    a tuning variable) changes the CTA directory's capacity: "6,10" leaves room for fewer than the 50 groups (the surplus
    keys live in the global table only), "8,3,2,2" and "2,2,6,3" run two and three CTAs per SM."""
    states = "ALAKAZARCACOCTDEFLGAHIIDILINIAKSKYLAMEMDMAMIMNMSMOMTNENVNHNJNMNYNCNDOHOKORPARISCSDTNTXUTVTVAWAWVWIWY"
    specs = [dict(kind=5, col_id=0, dict=states, dict_width=2), dict(kind=2, col_id=1, flo=0.0, fhi=1000.0, null_per_10k=100)]
    n = 300_000

    def run(E, batch):
        v = E.col(1)
        a = E.HashAggregate([E.col(0)], [("SUM", v), ("MIN", v), ("MAX", v), ("COUNT", v)], expected_groups=50) if E is G else \
            E.HashAggregate([E.col(0)], [("SUM", v), ("MIN", v), ("MAX", v), ("COUNT", v)])
        a.update(batch)
        return sorted(zip(*[x.to_pylist() for x in a.finalize().to_arrow()]))

    want = run(oracle, oracle.generate(specs, 5, 0, n))
    monkeypatch.setenv("KQ_AGG_GEOM", geom)
    got = run(G, G.generate(specs, 5, 0, n))
    assert len(got) == len(want) == 50
    for a, b in zip(got, want):
        assert a[0] == b[0] and a[2:] == b[2:] and abs(a[1] - b[1]) <= 1e-9 * abs(b[1])
