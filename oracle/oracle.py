"""ctypes wrapper over oracle/libkqoracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. It deliberately shares no code with the product binding
(query-engines_b200/kqgpu) so that a marshalling bug cannot cancel out on both sides.

Data goes in and out as pyarrow arrays (the same Arrow columnar layout as Arrow Java's
FieldVector: validity LSB-first, int32 offsets, raw bytes — SURVEY.md §7 step 3).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import pyarrow as pa

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkqoracle.so")

F64, UTF8, I64, BOOL, DATE32, I32 = 1, 2, 3, 4, 5, 6
OPS = {"EQ": 1, "NE": 2, "LT": 3, "LE": 4, "GT": 5, "GE": 6, "AND": 7, "OR": 8,
       "ADD": 9, "SUB": 10, "MUL": 11, "DIV": 12}
AGGS = {"MAX": 1, "MIN": 2, "SUM": 3, "COUNT": 4}
STATUS = {0: "OK", 1: "IllegalStateException", 2: "UnsupportedOperationException",
          3: "IllegalArgumentException", 4: "SQLException", 5: "NumberFormatException",
          6: "ArithmeticException"}


class GenSpec(C.Structure):
    """Mirror of kq_gen_spec (include/kq_gen.h)."""
    _fields_ = [("kind", C.c_int32), ("col_id", C.c_int32), ("ilo", C.c_int64), ("ihi", C.c_int64),
                ("flo", C.c_double), ("fhi", C.c_double), ("null_per_10k", C.c_int32),
                ("dict_width", C.c_int32), ("dict_count", C.c_int32), ("_pad", C.c_int32),
                ("dict", C.c_char_p)]


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


def build():
    """Compile the oracle if the shared object is missing or stale."""
    src = os.path.join(_HERE, "kq_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "kq_gen.h")
    if (not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libkqoracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ko_last_error.restype = C.c_char_p
        for name in ("ko_expr_column", "ko_expr_literal_f64", "ko_expr_literal_i64", "ko_expr_literal_bool",
                     "ko_expr_literal_date32", "ko_expr_literal_utf8", "ko_expr_literal_null",
                     "ko_expr_binary", "ko_expr_cast"):
            getattr(L, name).restype = C.c_void_p
        L.ko_expr_column.argtypes = [C.c_int]
        L.ko_expr_literal_f64.argtypes = [C.c_double]
        L.ko_expr_literal_i64.argtypes = [C.c_int64]
        L.ko_expr_literal_bool.argtypes = [C.c_int]
        L.ko_expr_literal_date32.argtypes = [C.c_int32]
        L.ko_expr_literal_utf8.argtypes = [C.c_char_p, C.c_int32]
        L.ko_expr_literal_null.argtypes = [C.c_int]
        L.ko_expr_binary.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.ko_expr_cast.argtypes = [C.c_void_p, C.c_int]
        L.ko_expr_free.argtypes = [C.c_void_p]
        L.ko_expr_free.restype = None
        L.ko_column_new.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.POINTER(C.c_void_p)]
        L.ko_column_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.ko_column_type.argtypes = [C.c_void_p]
        L.ko_column_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ko_column_free.argtypes = [C.c_void_p]
        L.ko_batch_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.POINTER(C.c_void_p)]
        L.ko_batch_num_rows.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.ko_batch_num_columns.argtypes = [C.c_void_p]
        L.ko_batch_column.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.ko_batch_free.argtypes = [C.c_void_p]
        L.ko_expr_evaluate.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.ko_project.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.ko_filter.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.ko_filter_project.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.ko_hashagg_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int),
                                        C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]
        L.ko_hashagg_update.argtypes = [C.c_void_p, C.c_void_p]
        L.ko_hashagg_finalize.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.ko_hashagg_free.argtypes = [C.c_void_p]
        L.ko_generate.argtypes = [C.POINTER(GenSpec), C.c_int, C.c_uint64, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]
        L.ko_csv_header.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.c_char_p]
        L.ko_csv_scan.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
        L.ko_filter_project_mt.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int,
                                           C.POINTER(C.c_int64)]
        L.ko_hashagg_mt.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int),
                                    C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.ko_parse_double.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_int)]
        L.ko_parse_double.restype = C.c_double
        _lib = L
    return _lib


def _check(code):
    if code != 0:
        raise OracleError(code, lib().ko_last_error().decode("utf-8", "replace"))


_PA_TYPES = {F64: pa.float64(), UTF8: pa.string(), I64: pa.int64(), BOOL: pa.bool_(),
             DATE32: pa.date32(), I32: pa.int32()}


def _type_of(arr: pa.Array) -> int:
    for k, t in _PA_TYPES.items():
        if arr.type == t:
            return k
    raise TypeError(f"unsupported arrow type {arr.type}")


def _addr(buf):
    return C.c_void_p(buf.address) if buf is not None and buf.size > 0 else C.c_void_p(0)


class Column:
    """ArrowFieldVector / ColumnVector (Main.kt:24-27, 176-202) on the oracle side."""

    def __init__(self, handle):
        self.h = handle

    @staticmethod
    def from_arrow(arr) -> "Column":
        if isinstance(arr, pa.ChunkedArray):
            arr = arr.combine_chunks()
        if arr.offset != 0:
            arr = pa.concat_arrays([arr])
        t = _type_of(arr)
        bufs = arr.buffers()
        validity = bufs[0] if arr.null_count > 0 else None
        out = C.c_void_p()
        if t == UTF8:
            data_bytes = bufs[2].size if bufs[2] is not None else 0
            _check(lib().ko_column_new(t, len(arr), _addr(validity), _addr(bufs[1]), _addr(bufs[2]), data_bytes, C.byref(out)))
        else:
            _check(lib().ko_column_new(t, len(arr), _addr(validity), None, _addr(bufs[1]), 0, C.byref(out)))
        return Column(out)

    def type(self):
        return lib().ko_column_type(self.h)

    def to_arrow(self) -> pa.Array:
        n, nb, nn = C.c_int64(), C.c_int64(), C.c_int64()
        _check(lib().ko_column_sizes(self.h, C.byref(n), C.byref(nb), C.byref(nn)))
        n, nb, nn = n.value, nb.value, nn.value
        t = self.type()
        validity = np.zeros((n + 7) // 8, dtype=np.uint8)
        data = np.zeros(max(nb, 1), dtype=np.uint8)
        offsets = np.zeros(n + 1, dtype=np.int32)
        _check(lib().ko_column_download(self.h, validity.ctypes.data, offsets.ctypes.data if t == UTF8 else None,
                                        data.ctypes.data))
        vbuf = pa.py_buffer(validity.tobytes()) if nn > 0 else None
        if t == UTF8:
            return pa.Array.from_buffers(pa.string(), n, [vbuf, pa.py_buffer(offsets.tobytes()),
                                                          pa.py_buffer(data[:nb].tobytes())], null_count=nn)
        return pa.Array.from_buffers(_PA_TYPES[t], n, [vbuf, pa.py_buffer(data[:nb].tobytes())], null_count=nn)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ko_column_free(self.h)
            self.h = None


class RecordBatch:
    """RecordBatch(schema, fields) (Main.kt:56-61)."""

    def __init__(self, handle):
        self.h = handle

    @staticmethod
    def from_arrow(arrays, n_rows=-1) -> "RecordBatch":
        cols = [Column.from_arrow(a) for a in arrays]
        return RecordBatch.from_columns(cols, n_rows)

    @staticmethod
    def from_columns(cols, n_rows=-1) -> "RecordBatch":
        arr = (C.c_void_p * max(len(cols), 1))(*[c.h for c in cols])
        out = C.c_void_p()
        _check(lib().ko_batch_create(arr, len(cols), n_rows, C.byref(out)))
        b = RecordBatch(out)
        b._keep = cols
        return b

    def row_count(self) -> int:
        n = C.c_int64()
        _check(lib().ko_batch_num_rows(self.h, C.byref(n)))
        return n.value

    def num_columns(self) -> int:
        return lib().ko_batch_num_columns(self.h)

    def field(self, i) -> Column:
        out = C.c_void_p()
        _check(lib().ko_batch_column(self.h, i, C.byref(out)))
        return Column(out)

    def to_arrow(self):
        return [self.field(i).to_arrow() for i in range(self.num_columns())]

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ko_batch_free(self.h)
            self.h = None


class Expr:
    def __init__(self, handle, keep=()):
        self.h = C.c_void_p(handle)
        self._keep = keep

    def evaluate(self, batch: RecordBatch) -> Column:
        """Expression.evaluate(input) (Main.kt:448-450)."""
        out = C.c_void_p()
        _check(lib().ko_expr_evaluate(self.h, batch.h, C.byref(out)))
        return Column(out)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ko_expr_free(self.h)
            self.h = None


def col(i): return Expr(lib().ko_expr_column(i))
def lit_f64(v): return Expr(lib().ko_expr_literal_f64(float(v)))
def lit_i64(v): return Expr(lib().ko_expr_literal_i64(int(v)))
def lit_bool(v): return Expr(lib().ko_expr_literal_bool(int(bool(v))))
def lit_date32(v): return Expr(lib().ko_expr_literal_date32(int(v)))
def lit_utf8(s):
    b = s.encode("utf-8") if isinstance(s, str) else bytes(s)
    return Expr(lib().ko_expr_literal_utf8(b, len(b)))
def lit_null(t): return Expr(lib().ko_expr_literal_null(t))
def binary(op, l, r): return Expr(lib().ko_expr_binary(OPS[op], l.h, r.h), (l, r))
def cast(e, t): return Expr(lib().ko_expr_cast(e.h, t), (e,))


def _expr_array(exprs):
    return (C.c_void_p * max(len(exprs), 1))(*[e.h for e in exprs])


def project(exprs, batch: RecordBatch) -> RecordBatch:
    """ProjectionExec.execute for one batch (Main.kt:589-594)."""
    out = C.c_void_p()
    _check(lib().ko_project(_expr_array(exprs), len(exprs), batch.h, C.byref(out)))
    return RecordBatch(out)


def filter(pred, batch: RecordBatch, want_selection=False):
    out, sel = C.c_void_p(), C.c_void_p()
    _check(lib().ko_filter(pred.h, batch.h, C.byref(out), C.byref(sel) if want_selection else None))
    if want_selection:
        return RecordBatch(out), Column(sel)
    return RecordBatch(out)


def filter_project(pred, exprs, batch: RecordBatch) -> RecordBatch:
    out = C.c_void_p()
    _check(lib().ko_filter_project(pred.h, _expr_array(exprs), len(exprs), batch.h, C.byref(out)))
    return RecordBatch(out)


class HashAggregate:
    """HashAggregateExec (Main.kt:605-660): update() per input batch, finalize() once."""

    def __init__(self, group_exprs, aggs, pred=None):
        # aggs: list of (kind_name, input_expr)
        self._keep = (group_exprs, aggs, pred)
        kinds = (C.c_int * max(len(aggs), 1))(*[AGGS[k] for k, _ in aggs])
        out = C.c_void_p()
        _check(lib().ko_hashagg_create(pred.h if pred else None, _expr_array(group_exprs), len(group_exprs), kinds,
                                       _expr_array([e for _, e in aggs]), len(aggs), C.byref(out)))
        self.h = out

    def update(self, batch: RecordBatch):
        _check(lib().ko_hashagg_update(self.h, batch.h))

    def finalize(self) -> RecordBatch:
        out = C.c_void_p()
        _check(lib().ko_hashagg_finalize(self.h, C.byref(out)))
        return RecordBatch(out)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ko_hashagg_free(self.h)
            self.h = None


def make_specs(specs):
    """specs: list of dicts with kq_gen_spec fields."""
    arr = (GenSpec * len(specs))()
    keep = []
    for i, s in enumerate(specs):
        g = arr[i]
        g.kind, g.col_id = s["kind"], s.get("col_id", i)
        g.ilo, g.ihi = s.get("ilo", 0), s.get("ihi", 1)
        g.flo, g.fhi = s.get("flo", 0.0), s.get("fhi", 1.0)
        g.null_per_10k = s.get("null_per_10k", 0)
        d = s.get("dict")
        if d is not None:
            d = d if isinstance(d, bytes) else d.encode()
            keep.append(d)
            g.dict = d
            g.dict_width = s["dict_width"]
            g.dict_count = len(d) // s["dict_width"]
    return arr, keep


def generate(specs, seed, row_begin, row_end) -> RecordBatch:
    arr, keep = make_specs(specs)
    out = C.c_void_p()
    _check(lib().ko_generate(arr, len(specs), seed, row_begin, row_end, C.byref(out)))
    return RecordBatch(out)


def csv_header(text: bytes, has_headers=True):
    """CsvDataSource.schema() (Main.kt:328-356): (column names, detected delimiter)."""
    names = C.create_string_buffer(1 << 16)
    n, d = C.c_int(), C.create_string_buffer(2)
    _check(lib().ko_csv_header(text, len(text), int(bool(has_headers)), names, len(names), C.byref(n), d))
    return names.value.decode("utf-8").split("\n")[:n.value], d.raw[:1].decode()


def csv_scan(text: bytes, has_headers=True, projection=None) -> RecordBatch:
    """CsvDataSource.scan(projection) (Main.kt:304-326) + createBatch (Main.kt:251-273); projection by column NAME."""
    idx = []
    if projection:
        names, _ = csv_header(text, has_headers)
        for p in projection:
            if p not in names:
                raise OracleError(3, f"Field {p} not found")        # E_ILLEGAL_ARGUMENT = IllegalArgumentException, Main.kt:49
            idx.append(names.index(p))
    arr = (C.c_int * max(len(idx), 1))(*idx)
    out = C.c_void_p()
    _check(lib().ko_csv_scan(text, len(text), int(bool(has_headers)), arr if idx else None, len(idx), C.byref(out)))
    return RecordBatch(out)


def filter_project_mt(pred, exprs, batch, nthreads) -> int:
    n = C.c_int64()
    _check(lib().ko_filter_project_mt(pred.h if pred else None, _expr_array(exprs), len(exprs), batch.h, nthreads, C.byref(n)))
    return n.value


def hashagg_mt(group_exprs, aggs, batch, nthreads, pred=None) -> RecordBatch:
    kinds = (C.c_int * max(len(aggs), 1))(*[AGGS[k] for k, _ in aggs])
    out = C.c_void_p()
    _check(lib().ko_hashagg_mt(pred.h if pred else None, _expr_array(group_exprs), len(group_exprs), kinds,
                               _expr_array([e for _, e in aggs]), len(aggs), batch.h, nthreads, C.byref(out)))
    return RecordBatch(out)


def parse_double(s: str) -> float:
    b = s.encode("utf-8")
    st = C.c_int()
    v = lib().ko_parse_double(b, len(b), C.byref(st))
    _check(st.value)
    return v
