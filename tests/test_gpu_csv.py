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
