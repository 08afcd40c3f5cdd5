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


# ---------------------------------------------------------------- randomised round trips through Python's csv.writer
from hypothesis import assume, given, settings, strategies as st

_ALPHABET = st.sampled_from(list("abcXYZ019 ,;|\t\"'\n\r-_äÅ日") )
_FIELD = st.text(_ALPHABET, max_size=8)


def _trim(v):            # String.trim(): leading/trailing chars <= U+0020 (Main.kt:263)
    a, b = 0, len(v)
    while a < b and v[a] <= " ":
        a += 1
    while b > a and v[b - 1] <= " ":
        b -= 1
    return v[a:b]


@settings(max_examples=300, deadline=None)
@given(ncols=st.integers(1, 5), data=st.data(), crlf=st.booleans())
def test_roundtrip_of_random_tables(oracle, ncols, data, crlf):
    """Whatever csv.writer emits for a table (quoting fields that hold delimiters, quotes or line breaks), the oracle reads
    back as the trimmed fields — with ',' forced as delimiter by a header that only contains commas."""
    rows = data.draw(st.lists(st.lists(_FIELD, min_size=ncols, max_size=ncols), max_size=6))
    header = [f"h{i}" for i in range(ncols)]
    buf = io.StringIO(newline="")
    w = csv.writer(buf, lineterminator="\r\n" if crlf else "\n", quoting=csv.QUOTE_MINIMAL)
    w.writerow(header)
    # a single empty column would be written as an empty LINE, which CsvDataSource skips (Main.kt:293): keep such rows out
    rows = [r for r in rows if not (ncols == 1 and r[0] == "")]
    # lone '\r' inside a field is only data when the file has '\n' terminators (rule C1); csv.writer quotes it either way
    w.writerows(rows)
    text = buf.getvalue().encode("utf-8")
    if ncols == 1:
        return                                          # no delimiter in the header: detection is free to pick any candidate
    got = columns(oracle.csv_scan(text, True))
    want = [[_trim(r[c]) for r in rows] for c in range(ncols)]
    assert got == want


# ---------------------------------------------------------------- the device algorithm, modelled in Python, against the oracle
import csv_device_model as dm


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_device_model_on_cases(oracle, case):
    _, text, hdr, names, want = case
    assert dm.scan(text, hdr) == want
    d, _ = dm.detect(text)
    assert chr(d) == oracle.csv_header(text, hdr)[1]


_LONG_FIELD = st.text(_ALPHABET, max_size=40)        # long enough for fields and records to straddle 64-byte blocks


@settings(max_examples=400, deadline=None)
@given(ncols=st.integers(2, 4), data=st.data(), crlf=st.booleans(), blanks=st.booleans())
def test_device_model_agrees_with_oracle_on_random_tables(oracle, ncols, data, crlf, blanks):
    rows = data.draw(st.lists(st.lists(_LONG_FIELD, min_size=ncols, max_size=ncols), max_size=8))
    eol = "\r\n" if crlf else "\n"
    buf = io.StringIO(newline="")
    w = csv.writer(buf, lineterminator=eol, quoting=csv.QUOTE_MINIMAL)
    w.writerow([f"h{i}" for i in range(ncols)])
    for r in rows:
        w.writerow(r)
        if blanks:
            buf.write(eol * data.draw(st.integers(0, 3)))       # empty lines between records (skipped, rule C3)
    text = buf.getvalue().encode("utf-8")
    assert dm.scan(text, True) == columns(oracle.csv_scan(text, True))


# ---------------------------------------------------------------- the reader (kq_csv_reader_*): pieces cut at record boundaries
def _concat(batches, ncols):
    out = [[] for _ in range(ncols)]
    for b in batches:
        for acc, col in zip(out, b):
            acc.extend(col)
    return out


@pytest.mark.parametrize("rows,crlf", [(1, False), (63, True), (2000, False), (2000, True)])
@pytest.mark.parametrize("piece", [96, 256, 1024, 1 << 20])
def test_reader_model_pieces_concatenate_to_the_whole_scan(oracle, rows, crlf, piece):
    """The reader's piece logic (cut after the last complete record, carry the tail, pad to 16 bytes with terminators,
    header skipped once, last piece terminated) on files with quoted delimiters, line breaks inside quotes, CRLF and
    empty lines: the batches concatenate to what the oracle reads from the whole text."""
    text = synthetic(rows, seed=rows + 1, crlf=crlf)
    want = columns(oracle.csv_scan(text, True))
    batches = list(dm.reader(text, True, piece))
    assert _concat(batches, len(want)) == want
    assert all(b[0] for b in batches)                              # batches without rows are not yielded (Main.kt:245-247)
    if piece >= len(text):
        assert len(batches) == 1


@settings(max_examples=200, deadline=None)
@given(ncols=st.integers(2, 4), data=st.data(), crlf=st.booleans(), blanks=st.booleans(), piece=st.sampled_from([64, 80, 128, 256]),
       hdr=st.booleans(), end_eol=st.booleans())
def test_reader_model_on_random_tables(oracle, ncols, data, crlf, blanks, piece, hdr, end_eol):
    rows = data.draw(st.lists(st.lists(st.text(_ALPHABET, max_size=12), min_size=ncols, max_size=ncols), max_size=12))
    eol = "\r\n" if crlf else "\n"
    buf = io.StringIO(newline="")
    w = csv.writer(buf, lineterminator=eol, quoting=csv.QUOTE_MINIMAL)
    w.writerow([f"h{i}" for i in range(ncols)])
    for r in rows:
        w.writerow(r)
        if blanks:
            buf.write(eol * data.draw(st.integers(0, 3)))
    text = buf.getvalue()
    if not end_eol:
        text = text.rstrip("\r\n")                                 # a last record without a line separator still ends
    text = text.encode("utf-8")
    want = columns(oracle.csv_scan(text, hdr))
    try:
        batches = list(dm.reader(text, hdr, piece))
    except OverflowError:
        assume(False)                                               # a record longer than the piece: the reader refuses (tested on the GPU)
    assert _concat(batches, len(want)) == (want if want and want[0] else [[] for _ in want])


@settings(max_examples=400, deadline=None)
@given(body=st.text(st.sampled_from(list('ab1 ,,;;\t|"""\n\n\r')), max_size=120), hdr=st.booleans(),
       lead=st.sampled_from(["", "h1,h2\n", 'x"y;z\n', '"h,1";h2\r\n', " \n\n", "a|b|c\n", "a\tb\n"]))
def test_product_header_detection_on_raw_texts(oracle, body, hdr, lead):
    """kq_csv_header of the shipped library (host code: terminator and delimiter detection, first record under rule C2) on
    arbitrary texts rich in quotes and all four candidate delimiters, against the oracle's tokenizer."""
    import kqgpu
    from oracle.oracle import OracleError
    text = (lead + body).encode("utf-8")
    try:
        want = oracle.csv_header(text, hdr)
    except OracleError:
        return                      # the text ends inside a quoted field: the scan reports it (the header call reads one record)
    assert kqgpu.Engine.csv_header(text, hdr) == want
