"""CPU tests that pin the oracle: golden vectors from the reference's only fixture, the semantics
rules R1-R12 / E1-E8, and agreement with the independent numpy restatement."""
import json
import math
import os

import numpy as np
import pyarrow as pa
import pytest

import np_ref
from planspec import b, build, col, lit, sort_rows

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "employee_golden.json")
F64, UTF8, I64, BOOL, DATE32 = 1, 2, 3, 4, 5


@pytest.fixture(scope="module")
def employee():
    with open(GOLDEN, encoding="utf-8") as f:
        return json.load(f)


def employee_batch(O, g):
    # CsvDataSource yields all-Utf8 columns in header order (Main.kt:345-348)
    return O.RecordBatch.from_arrow([pa.array(g["columns"][name], pa.string()) for name in g["schema"]])


# ---------------------------------------------------------------- golden vectors (employee.csv)
def test_golden_scan_roundtrip(oracle, employee):
    batch = employee_batch(oracle, employee)
    assert batch.row_count() == 3 and batch.num_columns() == 6
    out = batch.to_arrow()
    for name, arr in zip(employee["schema"], out):
        assert arr.to_pylist() == employee["columns"][name]
    # multi-byte UTF-8 survives the offsets/data path byte for byte
    last = out[employee["schema"].index("last_name")]
    assert last[2].as_py().encode("utf-8").hex() == employee["last_name_row3_utf8_hex"]


def test_golden_group_by_state_max_salary(oracle, employee):
    O = oracle
    batch = employee_batch(O, employee)
    s = employee["schema"]
    agg = O.HashAggregate([O.col(s.index("state"))], [("MAX", O.cast(O.col(s.index("salary")), F64))])
    agg.update(batch)
    keys, mx = agg.finalize().to_arrow()
    assert dict(zip(keys.to_pylist(), mx.to_pylist())) == employee["group_by_state_max_salary"]


@pytest.mark.parametrize("state,key", [("CO", "config1_where_state_eq_CO"), ("Uppsala", "where_state_eq_Uppsala")])
def test_golden_config1_filter_project(oracle, employee, state, key):
    O = oracle
    batch = employee_batch(O, employee)
    s = employee["schema"]
    names = ["id", "first_name", "last_name", "state", "salary"]
    pred = O.binary("EQ", O.col(s.index("state")), O.lit_utf8(state))
    out = O.filter_project(pred, [O.col(s.index(n)) for n in names], batch)
    got = {n: a.to_pylist() for n, a in zip(names, out.to_arrow())}
    assert got == employee[key]
    assert out.row_count() == len(employee[key]["id"])


# ---------------------------------------------------------------- rules from Main.kt
def test_R1_R2_null_and_value_model(oracle):
    O = oracle
    a = pa.array([1.5, None, -0.0, float("inf")], pa.float64())
    s = pa.array(["a", None, "", "Pärsson"], pa.string())
    out = O.RecordBatch.from_arrow([a, s]).to_arrow()
    assert out[0].to_pylist()[1] is None and out[1].to_pylist() == ["a", None, "", "Pärsson"]
    assert math.copysign(1, out[0][2].as_py()) == -1.0


def test_R4_column_expression_is_alias(oracle):
    O = oracle
    batch = O.RecordBatch.from_arrow([pa.array([1.0, 2.0]), pa.array(["x", "y"])])
    assert O.col(1).evaluate(batch).to_arrow().to_pylist() == ["x", "y"]
    with pytest.raises(O.OracleError) as e:
        O.col(5).evaluate(batch)
    assert e.value.code == 1


def test_R5_cast_utf8_to_double(oracle):
    O = oracle
    batch = O.RecordBatch.from_arrow([pa.array(["1337", " 6.5e1 ", None, "-0", "NaN", "Infinity", "0x1p3", "1d", ".5", "1."])])
    got = O.cast(O.col(0), F64).evaluate(batch).to_arrow().to_pylist()
    assert got[:4] == [1337.0, 65.0, None, -0.0] and math.copysign(1, got[3]) == -1
    assert math.isnan(got[4]) and got[5] == math.inf and got[6:] == [8.0, 1.0, 0.5, 1.0]
    for bad in ["", "abc", "1_0", "0x10", "nan", "inf", "1e", "--1", "1 2"]:
        with pytest.raises(O.OracleError) as e:
            O.cast(O.col(0), F64).evaluate(O.RecordBatch.from_arrow([pa.array([bad])]))
        assert e.value.code == 5, bad            # NumberFormatException
    with pytest.raises(O.OracleError) as e:      # cast to anything but Double (Main.kt:799)
        O.cast(O.col(0), UTF8).evaluate(batch)
    assert e.value.code == 1
    with pytest.raises(O.OracleError) as e:      # non-String source (Main.kt:792)
        O.cast(O.col(0), F64).evaluate(O.RecordBatch.from_arrow([pa.array([True])]))
    assert e.value.code == 1


def test_R7_group_key_equality(oracle):
    O = oracle
    k = pa.array([float("nan"), 0.0, -0.0, float("nan"), None, None, 1.0], pa.float64())
    v = pa.array([1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0], pa.float64())
    agg = O.HashAggregate([O.col(0)], [("COUNT", O.col(1)), ("MAX", O.col(1))])
    agg.update(O.RecordBatch.from_arrow([k, v]))
    keys, cnt, mx = [a.to_pylist() for a in agg.finalize().to_arrow()]
    got = {}
    for kk, c, m in zip(keys, cnt, mx):
        tag = "null" if kk is None else ("nan" if kk != kk else ("-0" if (kk == 0 and math.copysign(1, kk) < 0) else kk))
        got[tag] = (c, m)
    # NaN == NaN is one group; +0.0 and -0.0 are different groups; null is a legitimate group
    assert got == {"nan": (2, 4.0), 0.0: (1, 2.0), "-0": (1, 3.0), "null": (2, 6.0), 1.0: (1, 7.0)}


def test_R9_max_accumulator_nan_and_signed_zero_are_order_dependent(oracle):
    O = oracle
    def mx(vals, kind="MAX"):
        agg = O.HashAggregate([], [(kind, O.col(0))])
        agg.update(O.RecordBatch.from_arrow([pa.array(vals, pa.float64())]))
        return agg.finalize().to_arrow()[0].to_pylist()
    assert math.isnan(mx([float("nan"), 5.0])[0])          # a leading NaN sticks
    assert mx([5.0, float("nan")]) == [5.0]                # a later NaN is dropped
    assert math.copysign(1, mx([0.0, -0.0])[0]) == 1 and math.copysign(1, mx([-0.0, 0.0])[0]) == -1
    assert mx([None, None]) == [None]                      # all-null group => null
    assert mx([None, 3.0, None, 7.0, 7.0, -1.0]) == [7.0]
    assert mx([None, 3.0, None, 7.0, -1.0], "MIN") == [-1.0]


def test_R10_zero_rows_zero_groups_even_for_global_aggregate(oracle):
    O = oracle
    agg = O.HashAggregate([], [("MAX", O.col(0)), ("COUNT", O.col(0))])
    agg.update(O.RecordBatch.from_arrow([pa.array([], pa.float64())]))
    out = agg.finalize()
    assert out.row_count() == 0 and out.num_columns() == 2


def test_R11_batch_boundaries_do_not_matter(oracle):
    O = oracle
    rng = np.random.default_rng(7)
    k = rng.integers(0, 13, 5000)
    v = rng.integers(-1000, 1000, 5000).astype(np.float64)
    def run(splits):
        agg = O.HashAggregate([O.col(0)], [("SUM", O.col(1)), ("MIN", O.col(1)), ("MAX", O.col(1)), ("COUNT", O.col(1))])
        for a, bnd in zip(splits[:-1], splits[1:]):
            agg.update(O.RecordBatch.from_arrow([pa.array(k[a:bnd]), pa.array(v[a:bnd])]))
        return sort_rows(agg.finalize().to_arrow(), 1)
    assert run([0, 5000]) == run([0, 1, 1000, 1000, 4999, 5000])


def test_R12_count_is_int64_and_never_null(oracle):
    O = oracle
    agg = O.HashAggregate([O.col(0)], [("COUNT", O.col(1)), ("SUM", O.col(1))])
    agg.update(O.RecordBatch.from_arrow([pa.array(["a", "a", "b"]), pa.array([None, None, 2.0], pa.float64())]))
    keys, cnt, sm = agg.finalize().to_arrow()
    assert cnt.type == pa.int64()
    assert dict(zip(keys.to_pylist(), zip(cnt.to_pylist(), sm.to_pylist()))) == {"a": (0, None), "b": (1, 2.0)}


# ---------------------------------------------------------------- extension rules
def test_E2_operand_types_must_match(oracle):
    O = oracle
    batch = O.RecordBatch.from_arrow([pa.array([1, 2]), pa.array([1.0, 2.0])])
    with pytest.raises(O.OracleError) as e:
        O.binary("ADD", O.col(0), O.col(1)).evaluate(batch)
    assert e.value.code == 1
    with pytest.raises(O.OracleError):
        O.binary("AND", O.col(0), O.col(0)).evaluate(batch)


def test_E3_three_valued_logic_and_filter_drops_null(oracle):
    O = oracle
    T, F, N = True, False, None
    a = pa.array([T, T, T, F, F, F, N, N, N])
    c = pa.array([T, F, N, T, F, N, T, F, N])
    batch = O.RecordBatch.from_arrow([a, c])
    assert O.binary("AND", O.col(0), O.col(1)).evaluate(batch).to_arrow().to_pylist() == [T, F, N, F, F, F, N, F, N]
    assert O.binary("OR", O.col(0), O.col(1)).evaluate(batch).to_arrow().to_pylist() == [T, T, T, T, F, N, T, N, N]
    kept, sel = O.filter(O.binary("OR", O.col(0), O.col(1)), batch, want_selection=True)
    assert sel.to_arrow().to_pylist() == [0, 1, 2, 3, 6] and kept.row_count() == 5


def test_E4_float_math_is_not_fused_and_int_wraps(oracle):
    O = oracle
    # a*b+c where fma(a,b,c) != round(round(a*b)+c)
    a, bb, c = 1.0 + 2.0 ** -30, 1.0 + 2.0 ** -30, -(1.0 + 2.0 ** -29)
    batch = O.RecordBatch.from_arrow([pa.array([a]), pa.array([bb]), pa.array([c])])
    got = O.binary("ADD", O.binary("MUL", O.col(0), O.col(1)), O.col(2)).evaluate(batch).to_arrow()[0].as_py()
    assert got == (a * bb) + c and got != math.fma(a, bb, c) if hasattr(math, "fma") else True
    ib = O.RecordBatch.from_arrow([pa.array([2 ** 62, -2 ** 63]), pa.array([4, -1])])
    assert O.binary("MUL", O.col(0), O.col(1)).evaluate(ib).to_arrow().to_pylist() == [0, -2 ** 63]
    assert O.binary("DIV", O.col(0), O.col(1)).evaluate(ib).to_arrow().to_pylist() == [2 ** 60, -2 ** 63]
    with pytest.raises(O.OracleError) as e:
        O.binary("DIV", O.col(0), O.lit_i64(0)).evaluate(ib)
    assert e.value.code == 6


def test_utf8_comparisons_are_bytewise(oracle):
    O = oracle
    s = pa.array(["CO", "CA", "C", "COO", "", None, "Pärsson"])
    batch = O.RecordBatch.from_arrow([s])
    assert O.binary("EQ", O.col(0), O.lit_utf8("CO")).evaluate(batch).to_arrow().to_pylist() == [True, False, False, False, False, None, False]
    assert O.binary("LT", O.col(0), O.lit_utf8("CO")).evaluate(batch).to_arrow().to_pylist() == [False, True, True, False, True, None, False]
    assert O.binary("GE", O.col(0), O.lit_utf8("CO")).evaluate(batch).to_arrow().to_pylist() == [True, False, False, True, False, None, True]


# ---------------------------------------------------------------- oracle == numpy restatement
def rand_table(rng, n, null_frac):
    def mask(x):
        return pa.array(x, mask=rng.random(n) < null_frac) if null_frac else pa.array(x)
    return [mask(rng.integers(0, 1 << 20, n)), mask(rng.integers(0, 1 << 20, n)), mask(rng.integers(-50, 50, n)),
            mask(rng.random(n)), mask(rng.random(n)), mask(rng.random(n) * 1000),
            mask(np.array(["AL", "AK", "AZ", "CO", "NY", "Uppsala", ""], dtype=object)[rng.integers(0, 7, n)]),
            mask(rng.random(n) < 0.5)]


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
]


@pytest.mark.parametrize("null_frac", [0.0, 0.05])
@pytest.mark.parametrize("spec", SPECS, ids=[str(i) for i in range(len(SPECS))])
def test_expressions_match_numpy(oracle, spec, null_frac):
    rng = np.random.default_rng(42)
    n = 3001
    arrs = rand_table(rng, n, null_frac)
    got = build(oracle, spec).evaluate(oracle.RecordBatch.from_arrow(arrs)).to_arrow().to_pylist()
    want = np_ref.to_pylist(np_ref.evaluate(spec, [np_ref.from_arrow(a) for a in arrs], n))
    assert got == want


@pytest.mark.parametrize("null_frac", [0.0, 0.05])
def test_filter_project_matches_numpy(oracle, null_frac):
    rng = np.random.default_rng(1)
    n = 5000
    arrs = rand_table(rng, n, null_frac)
    pred, proj = SPECS[2], [SPECS[0], col(6), SPECS[1]]
    out = oracle.filter_project(build(oracle, pred), [build(oracle, p) for p in proj], oracle.RecordBatch.from_arrow(arrs))
    cols = [np_ref.from_arrow(a) for a in arrs]
    idx = np_ref.filter_rows(np_ref.evaluate(pred, cols, n))
    taken = [np_ref.take(c, idx) for c in cols]
    want = [np_ref.to_pylist(np_ref.evaluate(p, taken, len(idx))) for p in proj]
    assert [a.to_pylist() for a in out.to_arrow()] == want and out.row_count() == len(idx)


@pytest.mark.parametrize("null_frac", [0.0, 0.1])
@pytest.mark.parametrize("keys", [[6], [2], [6, 7], []])
def test_hash_aggregate_matches_numpy(oracle, keys, null_frac):
    O = oracle
    rng = np.random.default_rng(3)
    n = 4000
    arrs = rand_table(rng, n, null_frac)
    arrs[5] = pa.array(np.floor(arrs[5].fill_null(0).to_numpy(zero_copy_only=False)), mask=np.array(arrs[5].is_null()))  # exact sums
    aggs = [("SUM", 5), ("MIN", 5), ("MAX", 5), ("COUNT", 5), ("SUM", 2), ("MAX", 0)]
    agg = O.HashAggregate([O.col(k) for k in keys], [(kind, O.col(c)) for kind, c in aggs])
    agg.update(O.RecordBatch.from_arrow(arrs))
    out = agg.finalize().to_arrow()
    cols = [np_ref.from_arrow(a) for a in arrs]
    want = np_ref.group_aggregate([cols[k] for k in keys], [(kind, cols[c]) for kind, c in aggs])
    got = {tuple(r[:len(keys)]): list(r[len(keys):]) for r in zip(*[a.to_pylist() for a in out])}
    assert got == want


# ---------------------------------------------------------------- generator + partitioned execution
def gen_specs():
    return [dict(kind=1, ilo=0, ihi=1 << 20), dict(kind=2, flo=0.0, fhi=1000.0, null_per_10k=500),
            dict(kind=3, ilo=0, ihi=1000), dict(kind=4, ilo=0, ihi=11, fhi=100.0),
            dict(kind=5, dict="ALAKAZCO", dict_width=2), dict(kind=6, ilo=8036, ihi=10561), dict(kind=7, ilo=2500)]


def test_generator_is_a_pure_function_of_seed_column_row(oracle):
    O = oracle
    whole = O.generate(gen_specs(), 42, 0, 1000).to_arrow()
    parts = [O.generate(gen_specs(), 42, a, bnd).to_arrow() for a, bnd in [(0, 1), (1, 333), (333, 1000)]]
    for c in range(len(whole)):
        assert pa.concat_arrays([p[c] for p in parts]).to_pylist() == whole[c].to_pylist()
    assert whole[0].type == pa.int64() and 0 <= min(whole[0].to_pylist()) and max(whole[0].to_pylist()) < (1 << 20)
    assert 20 < whole[1].null_count < 90 and set(whole[4].to_pylist()) == {"AL", "AK", "AZ", "CO"}
    assert set(whole[3].to_pylist()) <= {i / 100.0 for i in range(11)}
    assert O.generate(gen_specs(), 43, 0, 1000).to_arrow()[0].to_pylist() != whole[0].to_pylist()


def test_partition_partial_merge_equals_single_pass(oracle):
    # the shape of main(): per-partition partial aggregate, then MAX-of-MAX etc. (Main.kt:1309-1325)
    O = oracle
    batch = O.generate(gen_specs(), 42, 0, 20000)
    aggs = [("SUM", O.col(2)), ("MIN", O.col(1)), ("MAX", O.col(1)), ("COUNT", O.col(1))]
    single = O.HashAggregate([O.col(4)], aggs)
    single.update(batch)
    merged = O.hashagg_mt([O.col(4)], aggs, batch, 4)
    assert sort_rows(merged.to_arrow(), 1) == sort_rows(single.finalize().to_arrow(), 1)
    pred = O.binary("GT", O.col(0), O.lit_i64(1 << 19))
    assert O.filter_project_mt(pred, [O.col(2)], batch, 3) == O.filter(pred, batch).row_count()
