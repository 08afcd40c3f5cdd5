"""csrc/kq_csv.cu compiled for the HOST (tests/host_shim/: kernels run as plain functions, one call per block and thread, in
sequence; device memory is host memory) and checked against the oracle: the real host orchestration of kq_csv_scan and of
the reader kq_csv_reader_open/next/close — format detection, projection, piece cutting, carried tails, alignment padding,
header skipping, the last piece's terminator, error paths — and the kernels' index arithmetic, where there is no GPU.
Stream ordering is not modelled (every call completes before it returns); the GPU suite covers that (tests/test_gpu_csv.py)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from csv_cases import CASES, synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "query-engines_b200", "csrc")
BUILD = os.path.join(HERE, "_build")


def _host_source():
    """kq_csv.cu with every `kernel<<<grid, block, smem, stream>>>(` rewritten to `KQ_LAUNCH(kernel, grid, block, `."""
    src = open(os.path.join(CSRC, "kq_csv.cu")).read()
    pat = re.compile(r"([A-Za-z_]\w*(?:<[^<>()]*>)?)<<<(.*?)>>>\(")
    def sub(m):
        cfg = [c.strip() for c in m.group(2).split(",")]
        assert len(cfg) == 4, m.group(0)
        return f"KQ_LAUNCH({m.group(1)}, {cfg[0]}, {cfg[1]}, "
    out, n = pat.subn(sub, src)
    assert n >= 6 and "<<<" not in out
    return out


@pytest.fixture(scope="module", params=[1024, 16], ids=["chunk1024", "chunk16"])
def H(request):
    """The host build; with 16 blocks per chunk (1 KiB of text) the chunk composition of the quote states, its vector
    loads and its tails are exercised by texts of a few KiB."""
    chunk = request.param
    os.environ["KQ_CSV_BATCH"] = "1" if chunk == 16 else "16"      # read once per library, at its first scan
    os.makedirs(BUILD, exist_ok=True)
    gen = os.path.join(BUILD, "kq_csv_host.cpp")
    with open(gen, "w") as f:
        f.write(_host_source())
    so = os.path.join(BUILD, f"libkqcsv_host_{chunk}.so")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas", "-Wno-subobject-linkage",
           f"-DKQ_CSV_CHUNK={chunk}",
           "-I", os.path.join(HERE, "host_shim"), "-I", CSRC, gen, os.path.join(HERE, "csv_host_harness.cpp"), "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    L = C.CDLL(so)
    P, PP = C.c_void_p, C.POINTER(C.c_void_p)
    L.kqh_ctx_new.restype = P
    L.kqh_ctx_free.argtypes = [P]
    L.kqh_device_buffer.restype = P
    L.kqh_device_buffer.argtypes = [P, C.c_char_p, C.c_int64, C.c_int]
    L.kqh_device_free.argtypes = [P, P]
    L.kq_last_error.restype = C.c_char_p
    L.kq_last_error.argtypes = [P]
    L.kq_csv_scan.argtypes = [P, P, C.c_int64, C.c_int, C.POINTER(C.c_int), C.c_int, PP]
    L.kq_csv_reader_open.argtypes = [P, P, C.c_int64, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int64, PP]
    L.kq_csv_reader_next.argtypes = [P, PP]
    L.kq_csv_reader_close.argtypes = [P]
    L.kq_batch_free.argtypes = [P]
    L.kqh_batch_rows.restype = C.c_int64
    L.kqh_batch_rows.argtypes = [P]
    L.kqh_batch_cols.argtypes = [P]
    L.kqh_col_bytes.restype = C.c_int64
    L.kqh_col_bytes.argtypes = [P, C.c_int]
    L.kqh_col_read.argtypes = [P, C.c_int, P, P]
    L.kqh_live_blocks.restype = C.c_int64
    h = Host(L)
    assert h.scan(b"a,b\n1,2\n") == [["1"], ["2"]]
    os.environ.pop("KQ_CSV_BATCH", None)
    return h


class HostError(Exception):
    pass


class Host:
    def __init__(self, L):
        self.L = L
        self.ctx = L.kqh_ctx_new()

    def check(self, st):
        if st != 0:
            raise HostError(f"{st}: {self.L.kq_last_error(self.ctx).decode()}")

    def read(self, b):
        """batch handle -> list of columns (lists of str); frees the batch"""
        L = self.L
        rows, cols = L.kqh_batch_rows(b), []
        for c in range(L.kqh_batch_cols(b)):
            nb = L.kqh_col_bytes(b, c)
            off = (C.c_int32 * (rows + 1))()
            data = C.create_string_buffer(max(nb, 1))
            L.kqh_col_read(b, c, off, data)
            assert off[0] == 0 and off[rows] == nb
            raw, o = data.raw, list(off)                            # one copy each (ctypes makes a new object on every access)
            cols.append([raw[o[i]:o[i + 1]].decode("utf-8") for i in range(rows)])
        L.kq_batch_free(b)
        return cols

    def _src(self, text, device, misalign):
        if device:
            p = self.L.kqh_device_buffer(self.ctx, text, len(text), misalign)
            return p, C.c_void_p(p + misalign)
        keep = C.create_string_buffer(text, max(len(text), 1))
        return keep, C.cast(keep, C.c_void_p)

    def scan(self, text, hdr=True, cols=None, device=False, misalign=0):
        keep, src = self._src(text, device, misalign)
        idx = list(cols or [])
        arr = (C.c_int * max(len(idx), 1))(*idx)
        out = C.c_void_p()
        try:
            self.check(self.L.kq_csv_scan(self.ctx, src, len(text), int(hdr), arr if idx else None, len(idx), C.byref(out)))
            return self.read(out)
        finally:
            if device:
                self.L.kqh_device_free(self.ctx, keep)

    def batches(self, text, hdr=True, cols=None, piece=0, device=False, misalign=0):
        keep, src = self._src(text, device, misalign)
        idx = list(cols or [])
        arr = (C.c_int * max(len(idx), 1))(*idx)
        rd = C.c_void_p()
        self.check(self.L.kq_csv_reader_open(self.ctx, src, len(text), int(hdr), arr if idx else None, len(idx), piece, C.byref(rd)))
        try:
            while True:
                out = C.c_void_p()
                self.check(self.L.kq_csv_reader_next(rd, C.byref(out)))
                if not out.value:
                    return
                yield self.read(out)
        finally:
            self.L.kq_csv_reader_close(rd)
            if device:
                self.L.kqh_device_free(self.ctx, keep)

    def clean(self):
        return self.L.kqh_live_blocks() == 0 and self.L.kqh_guard_errors() == 0


def columns(batch):
    return [a.to_pylist() for a in batch.to_arrow()]


def concat(batches, ncols):
    out = [[] for _ in range(ncols)]
    for b in batches:
        assert b and b[0]                                         # batches without rows are not yielded (Main.kt:245-247)
        for acc, col in zip(out, b):
            acc.extend(col)
    return out


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_scan_and_reader_on_the_cases(H, case):
    _, text, hdr, names, want = case
    assert H.scan(text, hdr) == want
    if text:
        assert H.scan(text, hdr, device=True) == want == H.scan(text, hdr, device=True, misalign=5)
    batches = list(H.batches(text, hdr))
    assert len(batches) == (1 if want and want[0] else 0)
    assert concat(batches, len(names)) == want
    assert H.clean()


@pytest.mark.parametrize("rows,crlf", [(1, False), (63, True), (3000, False), (3000, True)])
def test_synthetic_scan_matches_the_oracle(H, oracle, rows, crlf):
    text = synthetic(rows, seed=rows, crlf=crlf)
    want = columns(oracle.csv_scan(text, True))
    assert H.scan(text, True) == want
    assert H.scan(text, True, [4, 1, 4]) == [want[4], want[1], want[4]]
    pad = (-len(text)) % 16
    assert H.scan(text + (b"\r\n" if crlf else b"\n") * (pad // (2 if crlf else 1)), True, device=True) == want        # in place: resident, aligned, terminated
    assert H.clean()


@pytest.mark.parametrize("rows,crlf,piece", [(3000, False, 256), (3000, True, 256), (3000, True, 512), (3000, False, 4096), (3000, True, 1 << 16),
                                             (20_000, False, 8192)])
def test_reader_batches_concatenate_to_the_whole_scan(H, oracle, rows, crlf, piece):
    text = synthetic(rows, seed=rows + 1, crlf=crlf)
    want = columns(oracle.csv_scan(text, True))
    for device, misalign in ((False, 0), (True, 3)):
        batches = list(H.batches(text, True, piece=piece, device=device, misalign=misalign))
        assert len(batches) > 1 and sum(len(b[0]) for b in batches) == rows
        assert concat(batches, 6) == want
    stripped = text.rstrip(b"\r\n")                                 # a last record without a line separator still ends
    assert concat(list(H.batches(stripped, True, [5, 0], piece=piece)), 2) == [want[5], want[0]]
    assert concat(list(H.batches(text, False, piece=piece)), 6) == columns(oracle.csv_scan(text, False))
    assert H.clean()


def test_reader_errors_and_edges(H):
    long_record = b"a,b\n" + b"x" * 600 + b",1\n2,3\n"
    with pytest.raises(HostError, match="longer than the reader's piece"):
        list(H.batches(long_record, True, piece=256))
    assert concat(list(H.batches(long_record, True, piece=1024)), 2) == [["x" * 600, "2"], ["1", "3"]]
    text = synthetic(2000, seed=2) + b'7,"open,1,2,3,4\n'
    seen = 0
    with pytest.raises(HostError, match="quoted"):
        for b in H.batches(text, True, piece=4096):
            seen += len(b[0])
    assert 0 < seen <= 2000
    with pytest.raises(HostError, match="quoted"):
        H.scan(text, True)
    assert list(H.batches(b"", True)) == [] and list(H.batches(b"a,b\n", True)) == []
    assert list(H.batches(b"a,b\n1,2", True, cols=[1])) == [[["2"]]]
    with pytest.raises(HostError, match="out of range"):
        list(H.batches(b"a,b\n1,2\n", True, cols=[2]))
    assert H.clean()


def test_reader_model_and_host_build_cut_the_same_pieces(H):
    """The Python model of the reader (tests/csv_device_model.py) and the compiled host code yield the same batches."""
    import csv_device_model as dm
    text = synthetic(1500, seed=9, crlf=True)
    for piece in (256, 768, 2048):
        assert list(H.batches(text, True, piece=piece)) == list(dm.reader(text, True, piece))


# ---------------------------------------------------------------- rule C2 under fire: quotes anywhere, not only where a CSV writer puts them
from hypothesis import given, settings, strategies as st

_RAW = st.text(st.sampled_from(list('ab1 ,,,;\t"""\n\n\r|é')), max_size=200)


@settings(max_examples=600, deadline=None)
@given(body=_RAW, hdr=st.booleans(), lead=st.sampled_from(["h1,h2,h3\n", "h1;h2\r\n", "", 'x"y,z\n', '"h,1",h2\n']))
def test_raw_texts_with_stray_quotes_match_the_oracle(H, oracle, body, hdr, lead):
    """Arbitrary bytes from a small alphabet rich in quotes, separators, blanks and line breaks — texts no CSV writer
    would emit: quotes inside unquoted values, behind closed sections, doubled at block edges, unbalanced. The compiled
    device code, the Python model of it and the oracle's character-at-a-time tokenizer agree on every one, error or not."""
    import csv_device_model as dm
    from oracle.oracle import OracleError
    text = (lead + body).encode("utf-8")
    try:
        want = columns(oracle.csv_scan(text, hdr))
    except OracleError as e:
        assert "quoted" in str(e)
        with pytest.raises(HostError, match="quoted"):
            H.scan(text, hdr)
        with pytest.raises(ValueError, match="quoted"):
            dm.scan(text, hdr)
        return
    assert H.scan(text, hdr) == want
    model = dm.scan(text, hdr)
    assert model == want or (model == [] and all(c == [] for c in want))
    nonempty = want and want[0]
    for piece in (256, 512):
        try:
            got = list(H.batches(text, hdr, piece=piece))
        except HostError as e:
            assert "longer than the reader's piece" in str(e)
            continue
        assert concat(got, len(want)) == (want if nonempty else [[] for _ in want])
    assert H.clean()


@settings(max_examples=150, deadline=None)
@given(parts=st.lists(st.tuples(st.integers(0, 130), st.sampled_from(['"', '""', '" "', ',"', '"\n', '\n"', ' " ', '"x"', ',', '\n'])), max_size=12))
def test_quotes_at_every_block_offset(H, oracle, parts):
    """Runs of filler with a quote pattern dropped at arbitrary distances, so that patterns meet the 64-byte block edges
    (a doubled quote split over two blocks, an opener as a block's first byte behind blanks that end the block before)."""
    from oracle.oracle import OracleError
    text = "k,v\n" + "".join(("f" * (n % 7) + " " * (n // 7 % 5) + "," * (n % 2) + "g" * (n // 35)) + pat for n, pat in parts)
    text = text.encode()
    try:
        want = columns(oracle.csv_scan(text, True))
    except OracleError:
        with pytest.raises(HostError, match="quoted"):
            H.scan(text, True)
        return
    assert H.scan(text, True) == want
    assert H.scan(text, True, device=True, misalign=7) == want


def test_a_quoted_section_that_spans_many_blocks_and_chunks(H, oracle):
    """A 9 KiB quoted value full of delimiters, line breaks and doubled quotes, records in front of it and behind it:
    every block inside it is entered in the `inside` state, which only the composition of the blocks in front can tell."""
    inner = 'l,""i""\n' * 1100
    text = ('a,b\n' + '1,2\n' * 300 + '"' + inner + '",x\n' + '3,"4"\n' * 300 + 'it"s,"5\n6"\n').encode()
    want = columns(oracle.csv_scan(text, True))
    assert want[0][300] == inner.replace('""', '"').strip() and want[1][-1] == "5\n6" and want[0][-1] == 'it"s'
    assert H.scan(text, True) == want
    assert H.scan(text, True, device=True, misalign=9) == want
    assert concat(list(H.batches(text, True, piece=16384)), 2) == want
    with pytest.raises(HostError, match="longer than the reader's piece"):
        list(H.batches(text, True, piece=4096))
    assert H.clean()


def test_bench_text_at_reduced_size_is_periodic_like_its_blocks(H, oracle):
    """The GPU suite's full-size CSV check (tests/test_gpu_full_size.py: 10 M records) at a size the host build walks in
    seconds: bench.py's text of 250 000 records = a 100 000-record block repeated plus a cut block; the columns must be
    the oracle's columns of one block, repeated — here with every chunk of the quote-state composition several times over."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.WORKLOADS["csv"](250_000)
    text = wl.make_text(wl.rows)
    got = H.scan(text, True)
    block = columns(oracle.csv_scan(wl.make_text(100_000), True))
    tail = columns(oracle.csv_scan(wl.make_text(50_000), True))
    assert len(got) == 6 and len(got[0]) == 250_000
    for c in range(6):
        assert got[c][:100_000] == block[c] and got[c][100_000:200_000] == block[c] and got[c][200_000:] == tail[c]
    assert concat(list(H.batches(text, True, piece=1 << 20)), 6) == got
    assert H.clean()
