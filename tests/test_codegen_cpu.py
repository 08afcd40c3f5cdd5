"""CPU checks of the query compiler: expression trees -> CUDA source -> sm_100a cubin (NVRTC needs no GPU).

Covers the BASELINE.json query shapes, the type rules that must raise the reference's exception classes
(SURVEY.md §8b), and two properties of the generated text that parity depends on: Float64 math is emitted
as separately rounded __dmul_rn/__dadd_rn (the JVM never fuses a*b+c) and literal VALUES
are not baked into the source (one kernel per query shape).
"""
import pytest

F64, UTF8, I64, BOOL, D32 = 1, 2, 3, 4, 5


@pytest.fixture(scope="module")
def E():
    import build
    build.build()
    import kqgpu
    return kqgpu.Exprs()


def cfg2(E, lit):
    pred = E.binary("AND", E.binary("GT", E.col(0), lit), E.binary("LT", E.col(1), lit))
    proj = E.binary("ADD", E.binary("MUL", E.col(0), E.col(1)), E.col(2))
    return pred, [proj]


def test_filter_project_shapes_compile(E):
    pred, proj = cfg2(E, E.lit_f64(0.5))
    src = E.explain_filter_project(pred, proj, [F64] * 3)
    assert "__dmul_rn" in src and "__dadd_rn" in src and "fma" not in src.lower()
    assert "KQ_KERNEL_FILTER" in src and "load_valid" not in src          # non-nullable: no validity handling at all
    src_n = E.explain_filter_project(pred, proj, [F64] * 3, [1, 0, 1])
    assert "load_valid" in src_n
    pred_i, proj_i = cfg2(E, E.lit_i64(1 << 19))
    E.explain_filter_project(pred_i, proj_i, [I64] * 3)
    E.explain_filter_project(None, proj, [F64] * 3)                         # ProjectionExec alone
    # config 1: Utf8 equality + pass-through of a Date32 column and a cast
    E.explain_filter_project(E.binary("EQ", E.col(0), E.lit_utf8("CO")), [E.col(1), E.cast(E.col(0), F64)], [UTF8, D32], [1, 1])


def test_one_kernel_per_shape_not_per_literal(E):
    a = E.explain_filter_project(*cfg2(E, E.lit_f64(0.5)), [F64] * 3, compile=False)
    b = E.explain_filter_project(*cfg2(E, E.lit_f64(0.7)), [F64] * 3, compile=False)
    assert a == b


def test_aggregate_shapes_compile(E):
    v = E.col(1)
    four = [("SUM", v), ("MIN", v), ("MAX", v), ("COUNT", v)]
    src = E.explain_hashagg([E.col(0)], four, [UTF8, F64])                 # config 3
    assert "NCNT = 1" in src                                                # non-null input: COUNT shares the row counter
    E.explain_hashagg([E.col(0)], four, [I64, F64], [1, 1])                 # config 4, nullable
    one = E.lit_f64(1.0)
    dp = E.binary("MUL", E.col(4), E.binary("SUB", one, E.col(5)))
    ch = E.binary("MUL", dp, E.binary("ADD", one, E.col(6)))
    E.explain_hashagg([E.col(1), E.col(2)], [("SUM", E.col(3)), ("SUM", E.col(4)), ("SUM", dp), ("SUM", ch), ("COUNT", E.lit_i64(1))],
                      [D32, UTF8, UTF8, F64, F64, F64, F64], pred=E.binary("LE", E.col(0), E.lit_date32(10471)))   # config 5
    E.explain_hashagg([], [("MAX", v)], [UTF8, F64])                        # global aggregate (no group-by)
    E.explain_hashagg([E.col(0)], [("MAX", E.cast(E.col(1), F64))], [UTF8, UTF8])   # the reference's own query (Main.kt:1336)


def test_type_errors_raise_the_reference_exception_classes(E):
    import kqgpu
    with pytest.raises(kqgpu.KqError) as e:      # rule E2: operand types must match
        E.explain_filter_project(None, [E.binary("ADD", E.col(0), E.col(1))], [F64, I64])
    assert e.value.code == 1
    with pytest.raises(kqgpu.KqError) as e:      # Main.kt:799: only Double is a cast target
        E.explain_filter_project(None, [E.cast(E.col(0), I64)], [F64])
    assert e.value.code == 1
    with pytest.raises(kqgpu.KqError) as e:      # FilterExec needs a Bool predicate
        E.explain_filter_project(E.col(0), [E.col(0)], [F64])
    assert e.value.code == 1
    with pytest.raises(kqgpu.KqError) as e:      # MaxAccumulator: UnsupportedOperationException for other types (Main.kt:548)
        E.explain_hashagg([E.col(0)], [("SUM", E.col(1))], [I64, BOOL])
    assert e.value.code == 2
    with pytest.raises(kqgpu.KqError) as e:      # column index out of range: IllegalStateException
        E.explain_filter_project(None, [E.col(3)], [F64])
    assert e.value.code == 1


def _defines(src):
    import re
    return {k: int(v) for k, v in re.findall(r"#define KQ_(R|WARPS|STAGES|FE_GROUPS|DIR_SLOTS|CTAS|AGG_MODE) (\d+)", src)}


@pytest.mark.parametrize("hint,mode", [("6", 3), ("50", 3), ("64", 3), ("1000", 2), ("10000000", 0)])
def test_every_aggregate_kernel_family_compiles_per_cardinality_hint(E, hint, mode, monkeypatch):
    """The planner's group-count hint picks the kernel: <= 64 groups the CTA-directory kernel (kq_k_agg_fe.cuh, mode 3), a
    block-shared table up to ~1600 (mode 2), else the global-table kernel (mode 0; the partitioned variant below)."""
    monkeypatch.setenv("KQ_EXPLAIN_GROUPS", hint)
    if mode == 2:
        monkeypatch.setenv("KQ_EXPLAIN_PARTS", "0")          # kq_explain_hashagg: 0 selects the block-shared table variant
    v = E.col(1)
    src = E.explain_hashagg([E.col(0)], [("SUM", v), ("MIN", v), ("MAX", v), ("COUNT", v)], [UTF8, F64], [1, 1])
    d = _defines(src)
    assert d["AGG_MODE"] == mode
    if mode == 3:
        assert d["STAGES"] >= 2 and d["WARPS"] >= 2 and d["FE_GROUPS"] >= min(int(hint), 64) and d["DIR_SLOTS"] >= 8 * d["FE_GROUPS"]


def test_partitioned_kernels_compile(E, monkeypatch):
    monkeypatch.setenv("KQ_EXPLAIN_GROUPS", "10000000")
    monkeypatch.setenv("KQ_EXPLAIN_PARTS", "2048")
    v = E.col(1)
    src = E.explain_hashagg([E.col(0), E.col(2)], [("SUM", v), ("MIN", v), ("MAX", v), ("COUNT", v)], [I64, F64, BOOL], [1, 1, 1])
    assert _defines(src)["AGG_MODE"] == 1


@pytest.mark.parametrize("geom", ["8,7,2", "12,6,2", "4,4", "8,3,2,2", "2,2,6,3"])
def test_low_cardinality_kernel_compiles_for_the_geometries_the_gpu_tests_force(E, geom, monkeypatch):
    """KQ_AGG_GEOM = rows per thread, consumer warps[, stages[, CTAs per SM]] (tests/test_gpu_parity.py runs these on the
    device); with MIN/MAX the kernel also carries the housekeeping warp."""
    monkeypatch.setenv("KQ_EXPLAIN_GROUPS", "50")
    monkeypatch.setenv("KQ_AGG_GEOM", geom)
    v = E.col(1)
    want = [int(x) for x in geom.split(",")]
    for aggs in ([("SUM", v), ("MIN", v), ("MAX", v), ("COUNT", v)], [("SUM", v), ("COUNT", v)]):
        d = _defines(E.explain_hashagg([E.col(0)], aggs, [UTF8, F64]))
        assert (d["R"], d["WARPS"]) == (want[0], want[1])
        if len(want) > 3:
            assert d["CTAS"] == want[3]
