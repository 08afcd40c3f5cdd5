"""GPU parity of the CSV scan (kq_csv_scan = CsvDataSource.scan + createBatch, Main.kt:251-273, 276-357) against the
oracle, through the C ABI; BASELINE config 1 end to end from the bytes of the reference's employee.csv."""
import json
import os

import pytest

from csv_cases import CASES, synthetic

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "employee_golden.json")


@pytest.fixture(scope="module")
def G(gpu, gctx):
    return gpu.Engine(gctx)


def columns(batch):
    return [a.to_pylist() for a in batch.to_arrow()]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cases(G, oracle, case):
    _, text, hdr, names, want = case
    assert G.csv_header(text, hdr) == oracle.csv_header(text, hdr)
    assert G.csv_header(text, hdr)[0] == names
    b = G.csv_scan(text, hdr)
    assert b.row_count() == len(want[0]) and b.num_columns() == len(names)
    assert columns(b) == want == columns(oracle.csv_scan(text, hdr))


def test_empty_text(G):
    b = G.csv_scan(b"", True)
    assert b.num_columns() == 0 and b.row_count() == 0


def test_unterminated_quote_is_an_error(G, gpu):
    with pytest.raises(gpu.KqError, match="quoted"):
        G.csv_scan(b'a,b\n"open,2\n', True)


def test_projection(G, oracle, gpu):
    text = synthetic(3000)
    for sel in (["c3"], ["c5", "c0"], ["c1", "c1", "c4"]):
        assert columns(G.csv_scan(text, True, sel)) == columns(oracle.csv_scan(text, True, sel))
    with pytest.raises(gpu.KqError, match="not found"):
        G.csv_scan(text, True, ["nope"])


@pytest.mark.parametrize("rows,crlf", [(1, False), (63, True), (5000, False), (5000, True), (400_000, False)])
def test_synthetic_matches_oracle(G, oracle, rows, crlf):
    text = synthetic(rows, seed=rows, crlf=crlf)
    got, want = G.csv_scan(text, True), oracle.csv_scan(text, True)
    assert got.row_count() == want.row_count() == rows
    for a, b in zip(got.to_arrow(), want.to_arrow()):
        assert a.equals(b)


@pytest.mark.parametrize("state,key", [("CO", "config1_where_state_eq_CO"), ("Uppsala", "where_state_eq_Uppsala")])
def test_config1_end_to_end_from_csv_bytes(G, oracle, state, key):
    """BASELINE config 1: employee.csv via CsvDataSource: SELECT id, first_name, last_name, state, salary WHERE state = ..."""
    g = json.load(open(GOLDEN, encoding="utf-8"))
    text = bytes.fromhex(g["csv_text_hex"])
    names = G.csv_header(text, True)[0]
    assert names == g["schema"]
    sel = ["id", "first_name", "last_name", "state", "salary"]
    for E in (G, oracle):
        batch = E.csv_scan(text, True)
        pred = E.binary("EQ", E.col(names.index("state")), E.lit_utf8(state))
        out = E.filter_project(pred, [E.col(names.index(n)) for n in sel], batch)
        assert {n: a.to_pylist() for n, a in zip(sel, out.to_arrow())} == g[key]
        assert out.row_count() == len(g[key]["id"])


def test_scan_then_aggregate(G, oracle):
    """The shape of the reference's main() (Main.kt:1336): GROUP BY a Utf8 column, MAX(CAST(col AS double)) over a CSV scan."""
    text = synthetic(50_000, seed=3)
    rows = []
    for E in (G, oracle):
        batch = E.csv_scan(text, True, ["c1", "c2"])
        agg = E.HashAggregate([E.col(0)], [("MAX", E.cast(E.col(1), 1)), ("COUNT", E.col(1))])
        agg.update(batch)
        rows.append(sorted(zip(*[a.to_pylist() for a in agg.finalize().to_arrow()])))
    assert rows[0] == rows[1] and len(rows[0]) == 8


# ---------------------------------------------------------------- kq_csv_reader_*: the scan as a Sequence<RecordBatch>
def concat(batches, ncols):
    out = [[] for _ in range(ncols)]
    for b in batches:
        assert b.row_count() > 0                                   # batches without rows are not yielded (Main.kt:245-247)
        for acc, col in zip(out, columns(b)):
            acc.extend(col)
    return out


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reader_on_cases(G, case):
    _, text, hdr, names, want = case
    batches = list(G.csv_batches(text, hdr))
    assert len(batches) == (1 if want and want[0] else 0)
    assert concat(batches, len(names)) == want


@pytest.mark.parametrize("rows,crlf,piece", [(5000, False, 256), (5000, True, 512), (5000, True, 4096), (60_000, False, 1 << 16),
                                             (60_000, False, 0)])
def test_reader_batches_concatenate_to_the_whole_scan(G, oracle, rows, crlf, piece):
    """Pieces cut at record boundaries (quoted delimiters and line breaks, CRLF split across pieces, empty lines), the
    unfinished tail carried into the next piece: the batches concatenate to the oracle's scan of the whole text."""
    text = synthetic(rows, seed=rows + 1, crlf=crlf)
    want = columns(oracle.csv_scan(text, True))
    batches = list(G.csv_batches(text, True, piece_bytes=piece))
    assert (len(batches) > 4) if piece else (len(batches) == 1)
    assert sum(b.row_count() for b in batches) == rows
    assert concat(batches, 6) == want


def test_reader_projection_device_text_and_missing_final_terminator(G, oracle, gpu, gctx):
    import numpy as np
    import pyarrow as pa
    text = synthetic(3000, seed=5).rstrip(b"\n")                  # the last record ends with the text
    want = columns(oracle.csv_scan(text, True, ["c4", "c1", "c4"]))
    got = concat(list(G.csv_batches(text, True, ["c4", "c1", "c4"], piece_bytes=1024)), 3)
    assert got == want
    # the same text resident in HBM (a device pointer), at an odd address
    pad = (-(len(text) + 3)) % 8
    dev = gpu.Column.from_arrow(gctx, pa.array(np.frombuffer(b"\n\n\n" + text + b" " * pad, dtype=np.int64)))
    ptr = dev.device_ptrs()[2] + 3
    got = concat(list(G.csv_batches(ptr, True, nbytes=len(text), columns=[4, 1, 4], piece_bytes=2048)), 3)
    assert got == want
    # and through kq_csv_scan, in place when the resident text is aligned and terminated
    whole = synthetic(3000, seed=5)
    pad = (-len(whole)) % 8
    dev2 = gpu.Column.from_arrow(gctx, pa.array(np.frombuffer(whole + b"\n" * pad, dtype=np.int64)))
    assert columns(G.csv_scan_ptr(dev2.device_ptrs()[2], len(whole) + pad, True, [4, 1, 4])) == want


def test_reader_errors(G, gpu):
    long_record = b"a,b\n" + b"x" * 600 + b",1\n2,3\n"
    with pytest.raises(gpu.KqError, match="longer than the reader's piece"):
        list(G.csv_batches(long_record, True, piece_bytes=256))
    assert concat(list(G.csv_batches(long_record, True, piece_bytes=1024)), 2) == [["x" * 600, "2"], ["1", "3"]]
    # an unbalanced quote surfaces with the last piece, after the batches in front of it
    text = synthetic(2000, seed=2) + b'7,"open,1,2,3,4\n'
    it = G.csv_batches(text, True, piece_bytes=4096)
    with pytest.raises(gpu.KqError, match="quoted"):
        seen = 0
        for b in it:
            seen += b.row_count()
    assert 0 < seen <= 2000
    assert list(G.csv_batches(b"", True)) == [] and list(G.csv_batches(b"a,b\n", True)) == []
