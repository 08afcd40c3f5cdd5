"""Utf8 group keys of any length (Main.kt:621-627 builds the key from arbitrary Strings; `GROUP BY VendorID`, Main.kt:1336).

Keys of up to 7 bytes travel packed in one 64-bit word; longer ones are interned in the aggregate's key heap
(csrc/kq_rt.cuh utf8_intern): the key word is a hash of the bytes, the bytes are compared on every row, and finalize
reads them back from the heap. Checked against the CPU oracle, bit for bit on keys and integer-valued aggregates."""
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


def check(G, oracle, arrs, keys, aggs, hint=None, batches=None):
    kw = {} if hint is None else dict(expected_groups=hint)
    got, want = run(G, arrs, keys, aggs, batches, **kw), run(oracle, arrs, keys, aggs, batches)
    assert got.row_count() == want.row_count()
    assert sort_rows(got.to_arrow(), len(keys)) == sort_rows(want.to_arrow(), len(keys))


CITIES = ["Stockholm", "Uppsala", "Sthlm", "Göteborg", "Malmö", "", "San Francisco", "Llanfairpwllgwyngyllgogerychwyrndrobwllllantysiliogogogoch",
          "Å", "Örnsköldsvik", "x" * 8, "x" * 9, "x" * 40, "y" * 7, "日本語のキー", "München"]


@pytest.mark.parametrize("null_frac", [0.0, 0.08])
@pytest.mark.parametrize("n", [1, 33, 5000, 300_007])
def test_mixed_length_keys_match_the_oracle(G, oracle, n, null_frac):
    """7-, 8-, 9- and 40-byte keys, multi-byte UTF-8, the empty string, nulls; short and long keys in one column."""
    rng = np.random.default_rng(n)
    k = np.array(CITIES, dtype=object)[rng.integers(0, len(CITIES), n)]
    v = np.floor(rng.random(n) * 1000) - 300
    mask = rng.random(n) < null_frac if null_frac else None
    arrs = [pa.array(k, pa.string(), mask=mask), pa.array(v, pa.float64(), mask=(rng.random(n) < null_frac) if null_frac else None)]
    check(G, oracle, arrs, [0], [(a, 1) for a in FOUR], hint=len(CITIES) + 1)
    check(G, oracle, arrs, [0], [(a, 1) for a in FOUR])
    if n >= 5000:
        check(G, oracle, arrs, [0], [(a, 1) for a in FOUR], hint=20, batches=[0, n // 3, n // 3, n - 7, n])


def test_reference_query_shape_with_long_vendor_ids(G, oracle):
    """SELECT VendorID, MAX(CAST(fare_amount AS double)) ... GROUP BY VendorID (Main.kt:1336) with VendorIDs of 9+ bytes."""
    rng = np.random.default_rng(7)
    n = 60_000
    vendors = np.array(["Creative Mobile Technologies", "VeriFone Inc.", "CMT", "VTS-NEW-YORK"], dtype=object)[rng.integers(0, 4, n)]
    fares = np.array([f"{x:.2f}" for x in rng.random(n) * 80], dtype=object)
    arrs = [pa.array(vendors, pa.string()), pa.array(fares, pa.string())]
    def go(E, **kw):
        agg = E.HashAggregate([E.col(0)], [("MAX", E.cast(E.col(1), 1))], **kw)
        agg.update(E.RecordBatch.from_arrow(arrs))
        return sort_rows(agg.finalize().to_arrow(), 1)
    assert go(G, expected_groups=4) == go(oracle)


def test_many_distinct_long_keys_and_two_key_columns(G, oracle):
    """A few thousand distinct long keys (global-table path) and a long key next to a short one."""
    rng = np.random.default_rng(11)
    n = 120_000
    ids = rng.integers(0, 3000, n)
    k1 = np.array([f"customer-{i:06d}-{'z' * (i % 23)}" for i in range(3000)], dtype=object)[ids]
    k2 = np.array(["A", "N", "R", "return-flag-long"], dtype=object)[rng.integers(0, 4, n)]
    v = np.floor(rng.random(n) * 100)
    arrs = [pa.array(k1, pa.string()), pa.array(k2, pa.string()), pa.array(v, pa.float64())]
    check(G, oracle, arrs, [0], [(a, 2) for a in FOUR], hint=3000)
    check(G, oracle, arrs, [1, 0], [("SUM", 2), ("COUNT", 2)], hint=12_000, batches=[0, 50_000, n])
    check(G, oracle, arrs, [1], [(a, 2) for a in FOUR], hint=4)


def test_prefix_sharing_keys_are_distinct_groups(G, oracle):
    """Keys that share their first 7 bytes (what the old packed word kept) are different groups."""
    keys = ["Uppsala", "Uppsala ", "Uppsala1", "Uppsala12", "Uppsal", "Uppsalaä"]
    n = 6000
    k = pa.array([keys[i % len(keys)] for i in range(n)], pa.string())
    v = pa.array([float(i % 17) for i in range(n)], pa.float64())
    check(G, oracle, [k, v], [0], [(a, 1) for a in FOUR], hint=6)
