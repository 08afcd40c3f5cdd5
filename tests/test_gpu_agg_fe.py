"""Low-cardinality HashAggregateExec kernel (csrc/kq_k_agg_fe.cuh, Main.kt:605-660) against the CPU oracle, through the C ABI.

The cases aim at the kernel's own machinery rather than at the operator's semantics (tests/test_gpu_parity.py has those):
perfect placement with rebuilds (many keys, colliding keys), keys that do not fit the CTA directory (global-table
spill), a planner hint that is far off (optimistic table -> overflow -> rollback -> conservative sizing), the MIN/MAX
bounds (descending / ascending values, first values, ties, infinities, NaNs, Int64 extremes, all-null inputs) and
nullable keys/inputs across several update() calls.
"""
import math

import numpy as np
import pyarrow as pa
import pytest

from planspec import sort_rows

pytestmark = pytest.mark.gpu
FOUR = ("SUM", "MIN", "MAX", "COUNT")


@pytest.fixture(scope="module")
def G(gpu, gctx):
    return gpu.Engine(gctx)


def run(E, arrs, keys, aggs, batches=None, **kw):
    agg = E.HashAggregate([E.col(k) for k in keys], [(kind, E.col(c)) for kind, c in aggs], **kw)
    n = len(arrs[0])
    cuts = batches or [0, n]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        agg.update(E.RecordBatch.from_arrow([a.slice(lo, hi - lo) for a in arrs]))
    return agg.finalize()


def check(G, oracle, arrs, keys, aggs, hint=None, batches=None, sum_cols=()):
    kw = {} if hint is None else dict(expected_groups=hint)
    got, want = run(G, arrs, keys, aggs, batches, **kw), run(oracle, arrs, keys, aggs, batches)
    assert got.row_count() == want.row_count()
    g, w = sort_rows(got.to_arrow(), len(keys)), sort_rows(want.to_arrow(), len(keys))
    if not sum_cols:
        assert g == w
        return
    for a, b in zip(g, w):
        for i, (x, y) in enumerate(zip(a, b)):
            if i in sum_cols and x is not None and y is not None:
                assert abs(x - y) <= 1e-9 * max(abs(y), 1e-300), (a, b)
            else:
                assert x == y, (a, b)


def masked(rng, x, t, null_frac):
    return pa.array(x, type=t, mask=rng.random(len(x)) < null_frac) if null_frac else pa.array(x, type=t)


@pytest.mark.parametrize("ngroups", [1, 2, 7, 50, 64])
@pytest.mark.parametrize("null_frac", [0.0, 0.07])
def test_int_keys_fill_the_directory(G, oracle, ngroups, null_frac):
    """1..64 Int64 keys (64 = FE_MAX_GROUPS): every key must end in its home slot, whatever rebuilds that takes."""
    rng = np.random.default_rng(100 + ngroups)
    n = 200_003
    k = rng.integers(0, ngroups, n) * 7919 - 3
    v = np.floor(rng.random(n) * 1000) - 500
    arrs = [masked(rng, k, pa.int64(), null_frac), masked(rng, v, pa.float64(), null_frac)]
    check(G, oracle, arrs, [0], [(a, 1) for a in FOUR], hint=ngroups)
    check(G, oracle, arrs, [0], [(a, 1) for a in FOUR])                      # no hint: 64-group directory


def test_keys_that_collide_under_the_first_multipliers(G, oracle):
    """Keys chosen so that lo * s1 + hi * s2 (first multipliers 0x9E3779B1, 0x85EBCA6B) is the SAME 32-bit value:
    only a rebuild under other multipliers separates them."""
    s1, s2 = 0x9E3779B1, 0x85EBCA6B
    inv = pow(s1, -1, 1 << 32)
    keys = []
    for hi in range(1, 41):
        lo = (-(hi * s2) * inv) % (1 << 32)           # lo * s1 + hi * s2 == 0 (mod 2^32)
        assert (lo * s1 + hi * s2) % (1 << 32) == 0
        keys.append((hi << 32) | lo)
    rng = np.random.default_rng(7)
    n = 120_000
    k = np.array(keys, dtype=np.uint64)[rng.integers(0, len(keys), n)].astype(np.int64)
    v = np.floor(rng.random(n) * 100)
    check(G, oracle, [pa.array(k), pa.array(v)], [0], [(a, 1) for a in FOUR], hint=40)


@pytest.mark.parametrize("ngroups,hint", [(80, 50), (200, 8), (3000, 60)])
def test_more_groups_than_the_hint(G, oracle, ngroups, hint):
    """The hint is a sizing hint, not a limit: surplus keys go to the global table (atomics)."""
    rng = np.random.default_rng(ngroups)
    n = 150_000
    arrs = [pa.array(rng.integers(0, ngroups, n)), pa.array(np.floor(rng.random(n) * 1000))]
    check(G, oracle, arrs, [0], [(a, 1) for a in FOUR], hint=hint, batches=[0, 70_000, n])


def test_a_hint_that_is_far_off_rolls_back_and_grows(G, oracle):
    """400 k groups announced as 50: the optimistically sized table overflows, the launch is discarded, the table restored
    (it already holds the first batch's groups) and the batch redone with the conservative capacity rule."""
    rng = np.random.default_rng(5)
    n = 900_000
    k = np.concatenate([rng.integers(0, 40, 100_000), rng.integers(0, 400_000, n - 100_000)])
    arrs = [pa.array(k), pa.array(np.floor(rng.random(n) * 1000))]
    check(G, oracle, arrs, [0], [(a, 1) for a in FOUR], hint=50, batches=[0, 100_000, n])
    check(G, oracle, arrs, [0], [(a, 1) for a in FOUR], hint=None, batches=[0, n])


@pytest.mark.parametrize("order", ["descending", "ascending", "constant", "random"])
def test_min_max_bounds(G, oracle, order):
    """MIN/MAX only touch shared memory when a value beats the all-groups bound: descending values beat the MIN bound
    all the time, ascending ones the MAX bound, constants tie with it."""
    rng = np.random.default_rng(11)
    n = 400_000
    k = rng.integers(0, 50, n)
    base = {"descending": -np.arange(n, dtype=np.float64), "ascending": np.arange(n, dtype=np.float64),
            "constant": np.full(n, 42.0), "random": rng.standard_normal(n) * 1e6}[order]
    for t, v in ((pa.float64(), base), (pa.int64(), base.astype(np.int64))):
        arrs = [pa.array(k), masked(rng, v, t, 0.03)]
        check(G, oracle, arrs, [0], [("MIN", 1), ("MAX", 1), ("COUNT", 1)], hint=50)


def test_min_max_special_values(G, oracle):
    """Infinities as only values, NaNs after/between values, all-NaN groups, all-null groups, Int64 extremes."""
    nan, inf = float("nan"), float("inf")
    i64 = np.iinfo(np.int64)
    rows = [
        # key, f64 value, i64 value
        (0, inf, i64.max), (0, inf, i64.max),                 # MIN must be +inf / INT64_MAX (= the identities of the shared extremes)
        (1, -inf, i64.min), (1, -inf, i64.min),
        (2, 1.0, 5), (2, nan, 7), (2, -1.0, -7), (2, nan, 0),   # NaNs never replace a held value
        (3, nan, 1), (3, nan, 2),                             # all-NaN group -> NaN
        (4, None, None), (4, None, None),                     # all-null group -> null, COUNT 0
        (5, 7.0, 9), (5, None, None), (5, 3.0, -9),
        (6, -0.0, 0),
        (7, 1e308, 1), (7, -1e308, -1), (7, 5e-324, 0),
    ]
    reps = 3000                                                # many tiles, every warp sees every group
    k = pa.array([r[0] for r in rows] * reps, pa.int64())
    f = pa.array([r[1] for r in rows] * reps, pa.float64())
    i = pa.array([r[2] for r in rows] * reps, pa.int64())
    for E_hint in (8, None):
        kw = {} if E_hint is None else dict(expected_groups=E_hint)
        got = sort_rows(run(G, [k, f, i], [0], [("MIN", 1), ("MAX", 1), ("COUNT", 1), ("MIN", 2), ("MAX", 2), ("SUM", 2)], **kw).to_arrow(), 1)
        want = sort_rows(run(oracle, [k, f, i], [0], [("MIN", 1), ("MAX", 1), ("COUNT", 1), ("MIN", 2), ("MAX", 2), ("SUM", 2)]).to_arrow(), 1)
        assert len(got) == len(want) == 8
        for a, b in zip(got, want):
            for x, y in zip(a, b):
                assert (x is None and y is None) or (x is not None and y is not None and ((x != x and y != y) or x == y)), (a, b)


def test_first_value_arrives_late(G, oracle):
    """A group whose first non-null value shows up long after the bounds of the other groups are tight, and is neither
    below the MIN bound nor above the MAX bound: only the first-value rule gets it into the extremes."""
    rng = np.random.default_rng(3)
    n = 600_000
    k = rng.integers(0, 20, n)
    v = rng.random(n) * 1000 - 500                      # bounds settle near +-500 quickly
    late = np.flatnonzero(k == 7)
    vv = v.astype(object)
    cut = late[len(late) * 3 // 4]
    for j in late:
        vv[j] = None if j < cut else 0.125 + (j % 5)   # group 7: nulls for most of the table, then values inside every other group's range
    arrs = [pa.array(k), pa.array(list(vv), pa.float64())]
    check(G, oracle, arrs, [0], [(a, 1) for a in FOUR], hint=20, sum_cols=(1,))


@pytest.mark.parametrize("null_frac", [0.0, 0.1])
def test_multi_key_and_wide_aggregates(G, oracle, null_frac):
    """Utf8 + Bool + Date32 keys (three key words, nullable), six aggregate inputs, a few update() calls."""
    rng = np.random.default_rng(17)
    n = 90_000
    s = np.array(["A", "N", "R", "", "Uppsala", "Sthlm"], dtype=object)[rng.integers(0, 6, n)]
    arrs = [masked(rng, s, pa.string(), null_frac), masked(rng, rng.random(n) < 0.5, pa.bool_(), null_frac),
            masked(rng, rng.integers(9000, 9003, n).astype(np.int32), pa.int32(), null_frac).cast(pa.date32())]
    for j in range(6):
        arrs.append(masked(rng, np.floor(rng.random(n) * 100) - 50, pa.float64() if j % 2 else pa.int64(), null_frac))
    aggs = [("SUM", 3), ("MIN", 4), ("MAX", 5), ("COUNT", 6), ("SUM", 7), ("MAX", 8), ("MIN", 3), ("COUNT", 4)]       # six distinct inputs (the limit)
    check(G, oracle, arrs, [0, 1, 2], aggs, hint=64, batches=[0, 1, 30_000, 30_000, n])
    check(G, oracle, arrs, [2, 0], aggs, hint=20)
    check(G, oracle, arrs, [], aggs)


def test_filter_predicate_and_float_sums(G, oracle):
    rng = np.random.default_rng(23)
    n = 500_000
    arrs = [pa.array(rng.integers(0, 50, n)), pa.array(rng.random(n) * 1000), pa.array(rng.integers(0, 100, n))]
    def go(E, **kw):
        pred = E.binary("LT", E.col(2), E.lit_i64(37))
        agg = E.HashAggregate([E.col(0)], [(a, E.col(1)) for a in FOUR], pred=pred, **kw)
        agg.update(E.RecordBatch.from_arrow(arrs))
        return sort_rows(agg.finalize().to_arrow(), 1)
    got, want = go(G, expected_groups=50), go(oracle)
    assert len(got) == len(want) == 50
    for a, b in zip(got, want):
        assert a[0] == b[0] and a[2:] == b[2:] and abs(a[1] - b[1]) <= 1e-9 * abs(b[1])


def test_tiny_and_ragged_inputs(G, oracle):
    for n in (0, 1, 2, 31, 32, 33, 127, 128, 129, 1023, 1025, 4097):
        rng = np.random.default_rng(n)
        arrs = [pa.array(np.array(["AL", "AK", "", "Uppsala"], dtype=object)[rng.integers(0, 4, n)], pa.string()),
                pa.array(np.floor(rng.random(n) * 10))]
        check(G, oracle, arrs, [0], [(a, 1) for a in FOUR], hint=4)
