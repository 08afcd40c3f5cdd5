"""CPU: pins the oracle's restatement of CsvDataSource (Main.kt:276-357) — hand-written cases with spelled-out expected
values, the reference's own fixture (employee.csv), and an independent parser (Python's csv module) on synthetic files."""
import csv
import io
import json
import os

import pytest

from csv_cases import CASES, synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "employee_golden.json")


def columns(batch):
    return [a.to_pylist() for a in batch.to_arrow()]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cases(oracle, case):
    _, text, hdr, names, want = case
    assert oracle.csv_header(text, hdr)[0] == names
    b = oracle.csv_scan(text, hdr)
    assert columns(b) == want and b.row_count() == len(want[0])


def test_employee_fixture(oracle):
    g = json.load(open(GOLDEN, encoding="utf-8"))
    text = bytes.fromhex(g["csv_text_hex"])
    names, delim = oracle.csv_header(text, True)
    assert names == g["schema"] and delim == ","
    b = oracle.csv_scan(text, True)
    assert columns(b) == [g["columns"][n] for n in g["schema"]]            # all Utf8, trimmed (Main.kt:263)
    sel = ["id", "first_name", "last_name", "state", "salary"]              # BASELINE config 1's projection
    assert columns(oracle.csv_scan(text, True, sel)) == [g["columns"][n] for n in sel]
    with pytest.raises(Exception, match="not found"):                       # Schema.select, Main.kt:49
        oracle.csv_scan(text, True, ["nope"])


def test_unterminated_quote_is_an_error(oracle):
    with pytest.raises(Exception, match="quoted"):
        oracle.csv_scan(b'a,b\n"open,2\n', True)


@pytest.mark.parametrize("crlf", [False, True])
def test_agrees_with_python_csv_module(oracle, crlf):
    text = synthetic(5000, crlf=crlf)
    rows = [r for r in csv.reader(io.StringIO(text.decode("utf-8"), newline="")) if r]
    want = [[(r[c] if c < len(r) else "").strip() for r in rows[1:]] for c in range(len(rows[0]))]
    assert columns(oracle.csv_scan(text, True)) == want


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_product_header_detection_matches_oracle(oracle, case):
    """kq_csv_header is host code (format detection + first record): it runs without a GPU and must agree with the oracle."""
    import kqgpu
    _, text, hdr, names, _ = case
    assert kqgpu.Engine.csv_header(text, hdr) == oracle.csv_header(text, hdr)
    assert kqgpu.Engine.csv_header(text, hdr)[0] == names
