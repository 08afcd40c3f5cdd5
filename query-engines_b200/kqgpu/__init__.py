"""kqgpu — Python binding over the C ABI of libkqgpu.so (include/kqgpu.h).

This is plumbing for tests and bench.py: the product is the shared library, which a JVM host binds
through Panama FFM / JNI. There is NO CPU fallback: importing works anywhere (so
the symbol table can be checked on a CPU box) but creating a Context without a CUDA device raises.

`Engine(ctx)` exposes the same vocabulary as the reference's physical layer: col(i) is
ColumnExpression (Main.kt:452-460), cast is CastExpression (772-805), project is
ProjectionExec.execute for one batch (589-594), HashAggregate is HashAggregateExec (605-660);
literals, binary expressions, filter and SUM/MIN/COUNT are the extensions of SURVEY.md §8 a12.
The layer above — DataFrame, logical plan, planner, SQL — is in the submodules `plan` and `sql`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import pyarrow as pa

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_ROOT, "libkqgpu.so")

F64, UTF8, I64, BOOL, DATE32, I32 = 1, 2, 3, 4, 5, 6
OPS = {"EQ": 1, "NE": 2, "LT": 3, "LE": 4, "GT": 5, "GE": 6, "AND": 7, "OR": 8,
       "ADD": 9, "SUB": 10, "MUL": 11, "DIV": 12}
AGGS = {"MAX": 1, "MIN": 2, "SUM": 3, "COUNT": 4}
COMM_ID_BYTES = 128

_PA_TYPES = {F64: pa.float64(), UTF8: pa.string(), I64: pa.int64(), BOOL: pa.bool_(),
             DATE32: pa.date32(), I32: pa.int32()}


class GenSpec(C.Structure):
    """kq_gen_spec (include/kq_gen.h)."""
    _fields_ = [("kind", C.c_int32), ("col_id", C.c_int32), ("ilo", C.c_int64), ("ihi", C.c_int64),
                ("flo", C.c_double), ("fhi", C.c_double), ("null_per_10k", C.c_int32),
                ("dict_width", C.c_int32), ("dict_count", C.c_int32), ("_pad", C.c_int32),
                ("dict", C.c_char_p)]


class KqError(Exception):
    """Carries the kq_status code; .exception_class is the reference exception it stands for."""

    def __init__(self, code, msg):
        self.code = code
        self.exception_class = lib().kq_status_name(code).decode()
        super().__init__(f"{self.exception_class}: {msg}")


# every exported symbol of include/kqgpu.h: (restype, argtypes)
_P, _PP = C.c_void_p, C.POINTER(C.c_void_p)
_I64P = C.POINTER(C.c_int64)
SYMBOLS = {
    "kq_version": (C.c_char_p, []),
    "kq_status_name": (C.c_char_p, [C.c_int]),
    "kq_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "kq_ctx_create": (C.c_int, [C.c_int, _PP]),
    "kq_ctx_destroy": (C.c_int, [_P]),
    "kq_last_error": (C.c_char_p, [_P]),
    "kq_ctx_sync": (C.c_int, [_P]),
    "kq_ctx_stream": (C.c_void_p, [_P]),
    "kq_ctx_launch_count": (C.c_int64, [_P]),
    "kq_timer_begin": (C.c_int, [_P]),
    "kq_timer_end": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "kq_flush_l2": (C.c_int, [_P, C.c_size_t]),
    "kq_host_alloc": (C.c_int, [_P, C.c_size_t, _PP]),
    "kq_host_free": (C.c_int, [_P, _P]),
    "kq_host_register": (C.c_int, [_P, _P, C.c_size_t]),
    "kq_host_unregister": (C.c_int, [_P, _P]),
    "kq_column_upload": (C.c_int, [_P, C.c_int, C.c_int64, _P, _P, _P, C.c_int64, _PP]),
    "kq_column_sizes": (C.c_int, [_P, _P, _I64P, _I64P, _I64P]),
    "kq_column_type": (C.c_int, [_P]),
    "kq_column_download": (C.c_int, [_P, _P, _P, _P, _P]),
    "kq_column_device_ptrs": (C.c_int, [_P, _P, _PP, _PP, _PP]),
    "kq_column_retain": (C.c_int, [_P]),
    "kq_column_free": (C.c_int, [_P]),
    "kq_batch_create": (C.c_int, [_P, _PP, C.c_int, C.c_int64, _PP]),
    "kq_batch_num_rows": (C.c_int, [_P, _P, _I64P]),
    "kq_batch_num_columns": (C.c_int, [_P]),
    "kq_batch_column": (C.c_int, [_P, _P, C.c_int, _PP]),
    "kq_batch_free": (C.c_int, [_P]),
    "kq_expr_column": (C.c_void_p, [C.c_int]),
    "kq_expr_literal_f64": (C.c_void_p, [C.c_double]),
    "kq_expr_literal_i64": (C.c_void_p, [C.c_int64]),
    "kq_expr_literal_bool": (C.c_void_p, [C.c_int]),
    "kq_expr_literal_date32": (C.c_void_p, [C.c_int32]),
    "kq_expr_literal_utf8": (C.c_void_p, [C.c_char_p, C.c_int32]),
    "kq_expr_literal_null": (C.c_void_p, [C.c_int]),
    "kq_expr_binary": (C.c_void_p, [C.c_int, _P, _P]),
    "kq_expr_cast": (C.c_void_p, [_P, C.c_int]),
    "kq_expr_free": (None, [_P]),
    "kq_expr_evaluate": (C.c_int, [_P, _P, _P, _PP]),
    "kq_project": (C.c_int, [_P, _PP, C.c_int, _P, _PP]),
    "kq_filter": (C.c_int, [_P, _P, _P, _PP, _PP]),
    "kq_filter_project": (C.c_int, [_P, _P, _PP, C.c_int, _P, _PP]),
    "kq_filter_project_host": (C.c_int, [_P, _P, _PP, C.c_int, C.c_int, C.POINTER(C.c_int), _PP, _PP, C.c_int64,
                                          _PP, _PP, _I64P]),
    "kq_explain_filter_project": (C.c_int, [_P, _PP, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                                             C.c_char_p, C.c_size_t]),
    "kq_explain_hashagg": (C.c_int, [_P, _PP, C.c_int, C.POINTER(C.c_int), _PP, C.c_int, C.c_int, C.POINTER(C.c_int),
                                      C.POINTER(C.c_int), C.c_int, C.c_char_p, C.c_size_t]),
    "kq_hashagg_create": (C.c_int, [_P, _P, _PP, C.c_int, C.POINTER(C.c_int), _PP, C.c_int, C.c_int64, _PP]),
    "kq_hashagg_update": (C.c_int, [_P, _P, _P]),
    "kq_hashagg_finalize": (C.c_int, [_P, _P, _PP]),
    "kq_hashagg_num_groups": (C.c_int, [_P, _P, _I64P]),
    "kq_hashagg_free": (C.c_int, [_P]),
    "kq_comm_unique_id": (C.c_int, [_P, C.c_char_p]),
    "kq_comm_init": (C.c_int, [_P, C.c_char_p, C.c_int, C.c_int]),
    "kq_comm_destroy": (C.c_int, [_P]),
    "kq_comm_barrier": (C.c_int, [_P]),
    "kq_comm_allreduce_max_f32": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "kq_hashagg_merge_allreduce": (C.c_int, [_P, _P]),
    "kq_hashagg_repartition_alltoall": (C.c_int, [_P, _P]),
    "kq_generate": (C.c_int, [_P, C.POINTER(GenSpec), C.c_int, C.c_uint64, C.c_int64, C.c_int64, _PP]),
    "kq_csv_header": (C.c_int, [C.c_char_p, C.c_int64, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.c_char_p]),
    "kq_csv_scan": (C.c_int, [_P, C.c_char_p, C.c_int64, C.c_int, C.POINTER(C.c_int), C.c_int, _PP]),
    "kq_csv_reader_open": (C.c_int, [_P, C.c_char_p, C.c_int64, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int64, _PP]),
    "kq_csv_reader_next": (C.c_int, [_P, _PP]),
    "kq_csv_reader_close": (C.c_int, [_P]),
}

_lib = None


def lib():
    """Load libkqgpu.so. Fails loudly if the CUDA extension has not been built — never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python query-engines_b200/build.py` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def device_count() -> int:
    n = C.c_int(0)
    lib().kq_device_count(C.byref(n))
    return n.value


def _addr(buf):
    return C.c_void_p(buf.address) if buf is not None and buf.size > 0 else C.c_void_p(0)


def _type_of(arr: pa.Array) -> int:
    for k, t in _PA_TYPES.items():
        if arr.type == t:
            return k
    raise TypeError(f"unsupported arrow type {arr.type}")


class Context:
    """kq_ctx: one device, one compute stream, side copy streams. One per plan/thread."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        st = lib().kq_ctx_create(device, C.byref(h))
        if st != 0:
            raise KqError(st, "kq_ctx_create failed (no CUDA device? there is no CPU fallback)")
        self.h = h
        self.device = device

    def check(self, st):
        if st != 0:
            raise KqError(st, lib().kq_last_error(self.h).decode("utf-8", "replace"))

    def sync(self):
        self.check(lib().kq_ctx_sync(self.h))

    def launch_count(self) -> int:
        return lib().kq_ctx_launch_count(self.h)

    def timer_begin(self):
        self.check(lib().kq_timer_begin(self.h))

    def timer_end(self) -> float:
        ms = C.c_float()
        self.check(lib().kq_timer_end(self.h, C.byref(ms)))
        return ms.value

    def flush_l2(self, nbytes=256 << 20):
        self.check(lib().kq_flush_l2(self.h, nbytes))

    def host_alloc(self, nbytes) -> int:
        p = C.c_void_p()
        self.check(lib().kq_host_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def host_free(self, ptr):
        self.check(lib().kq_host_free(self.h, C.c_void_p(ptr)))

    def close(self):
        if self.h:
            lib().kq_ctx_destroy(self.h)
            self.h = None


class Column:
    """Device-resident ColumnVector (ArrowFieldVector, Main.kt:176-202). getValue(i) downloads."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.h = ctx, handle
        self._host = None

    @staticmethod
    def from_arrow(ctx: Context, arr) -> "Column":
        if isinstance(arr, pa.ChunkedArray):
            arr = arr.combine_chunks()
        if arr.offset != 0:
            arr = pa.concat_arrays([arr])
        t = _type_of(arr)
        bufs = arr.buffers()
        validity = bufs[0] if arr.null_count > 0 else None
        out = C.c_void_p()
        if t == UTF8:
            nbytes = bufs[2].size if bufs[2] is not None else 0
            ctx.check(lib().kq_column_upload(ctx.h, t, len(arr), _addr(validity), _addr(bufs[1]), _addr(bufs[2]), nbytes, C.byref(out)))
        else:
            ctx.check(lib().kq_column_upload(ctx.h, t, len(arr), _addr(validity), None, _addr(bufs[1]), 0, C.byref(out)))
        return Column(ctx, out)

    def type(self) -> int:
        return lib().kq_column_type(self.h)

    def sizes(self):
        n, nb, nn = C.c_int64(), C.c_int64(), C.c_int64()
        self.ctx.check(lib().kq_column_sizes(self.ctx.h, self.h, C.byref(n), C.byref(nb), C.byref(nn)))
        return n.value, nb.value, nn.value

    def size(self) -> int:
        n = C.c_int64()
        self.ctx.check(lib().kq_column_sizes(self.ctx.h, self.h, C.byref(n), None, None))
        return n.value

    def device_ptrs(self):
        v, o, d = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self.ctx.check(lib().kq_column_device_ptrs(self.ctx.h, self.h, C.byref(v), C.byref(o), C.byref(d)))
        return v.value, o.value, d.value

    def to_arrow(self) -> pa.Array:
        n, nb, nn = self.sizes()
        t = self.type()
        validity = np.zeros((n + 7) // 8 + 8, dtype=np.uint8)
        data = np.zeros(max(nb, 1) + 8, dtype=np.uint8)
        offsets = np.zeros(n + 1, dtype=np.int32)
        self.ctx.check(lib().kq_column_download(self.ctx.h, self.h, validity.ctypes.data,
                                                 offsets.ctypes.data if t == UTF8 else None, data.ctypes.data))
        vbuf = pa.py_buffer(validity[:(n + 7) // 8].tobytes()) if nn > 0 else None
        if t == UTF8:
            return pa.Array.from_buffers(pa.string(), n, [vbuf, pa.py_buffer(offsets.tobytes()), pa.py_buffer(data[:nb].tobytes())], null_count=nn)
        return pa.Array.from_buffers(_PA_TYPES[t], n, [vbuf, pa.py_buffer(data[:nb].tobytes())], null_count=nn)

    def get_value(self, i):
        """ColumnVector.getValue(i): Any? (Main.kt:178-197) — lazy download, then host access."""
        if self._host is None:
            self._host = self.to_arrow()
        return self._host[i].as_py()

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None and getattr(self.ctx, "h", None):
            _lib.kq_column_free(self.h)
            self.h = None


class RecordBatch:
    """RecordBatch(schema, fields) (Main.kt:56-61) over device columns."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.h = ctx, handle

    @staticmethod
    def from_columns(ctx, cols, n_rows=-1) -> "RecordBatch":
        arr = (C.c_void_p * max(len(cols), 1))(*[c.h for c in cols])
        out = C.c_void_p()
        ctx.check(lib().kq_batch_create(ctx.h, arr, len(cols), n_rows, C.byref(out)))
        return RecordBatch(ctx, out)

    def row_count(self) -> int:
        n = C.c_int64()
        self.ctx.check(lib().kq_batch_num_rows(self.ctx.h, self.h, C.byref(n)))
        return n.value

    def num_columns(self) -> int:
        return lib().kq_batch_num_columns(self.h)

    def field(self, i) -> Column:
        out = C.c_void_p()
        self.ctx.check(lib().kq_batch_column(self.ctx.h, self.h, i, C.byref(out)))
        return Column(self.ctx, out)

    def to_arrow(self):
        return [self.field(i).to_arrow() for i in range(self.num_columns())]

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None and getattr(self.ctx, "h", None):
            _lib.kq_batch_free(self.h)
            self.h = None


class Expr:
    """Expression (Main.kt:448-450). Pure host tree; compiled into one fused kernel at evaluate time."""

    def __init__(self, engine, handle, keep=()):
        if not handle:
            raise ValueError("null expression handle")
        self.engine, self.h, self._keep = engine, C.c_void_p(handle), keep

    def evaluate(self, batch: RecordBatch) -> Column:
        ctx = self.engine.ctx
        out = C.c_void_p()
        ctx.check(lib().kq_expr_evaluate(ctx.h, self.h, batch.h, C.byref(out)))
        return Column(ctx, out)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.kq_expr_free(self.h)
            self.h = None


def _expr_array(exprs):
    return (C.c_void_p * max(len(exprs), 1))(*[e.h for e in exprs])


class HashAggregate:
    """HashAggregateExec (Main.kt:605-660): update() per input batch, finalize() emits one batch."""

    def __init__(self, engine, group_exprs, aggs, pred=None, expected_groups=0):
        self.engine, self.ctx = engine, engine.ctx
        self._keep = (group_exprs, aggs, pred)
        kinds = (C.c_int * max(len(aggs), 1))(*[AGGS[k] for k, _ in aggs])
        out = C.c_void_p()
        self.ctx.check(lib().kq_hashagg_create(self.ctx.h, pred.h if pred else None, _expr_array(group_exprs), len(group_exprs),
                                               kinds, _expr_array([e for _, e in aggs]), len(aggs), expected_groups, C.byref(out)))
        self.h = out

    def update(self, batch: RecordBatch):
        self.ctx.check(lib().kq_hashagg_update(self.ctx.h, self.h, batch.h))

    def finalize(self) -> RecordBatch:
        out = C.c_void_p()
        self.ctx.check(lib().kq_hashagg_finalize(self.ctx.h, self.h, C.byref(out)))
        return RecordBatch(self.ctx, out)

    def num_groups(self) -> int:
        n = C.c_int64()
        self.ctx.check(lib().kq_hashagg_num_groups(self.ctx.h, self.h, C.byref(n)))
        return n.value

    def merge_allreduce(self):
        self.ctx.check(lib().kq_hashagg_merge_allreduce(self.ctx.h, self.h))

    def repartition_alltoall(self):
        self.ctx.check(lib().kq_hashagg_repartition_alltoall(self.ctx.h, self.h))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None and getattr(self.ctx, "h", None):
            _lib.kq_hashagg_free(self.h)
            self.h = None


def make_specs(specs):
    """specs: list of dicts with kq_gen_spec fields -> (ctypes array, keep-alive list)."""
    arr = (GenSpec * max(len(specs), 1))()
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


class Exprs:
    """Expression factory (pure host objects: works without a device). Mirrors the reference's physical
    expressions: col = ColumnExpression (Main.kt:452-460), cast = CastExpression (Main.kt:772-805),
    lit_* / binary = the Literal/Binary extensions."""

    # expressions
    def col(self, i): return Expr(self, lib().kq_expr_column(i))
    def lit_f64(self, v): return Expr(self, lib().kq_expr_literal_f64(float(v)))
    def lit_i64(self, v): return Expr(self, lib().kq_expr_literal_i64(int(v)))
    def lit_bool(self, v): return Expr(self, lib().kq_expr_literal_bool(int(bool(v))))
    def lit_date32(self, v): return Expr(self, lib().kq_expr_literal_date32(int(v)))

    def lit_utf8(self, s):
        b = s.encode("utf-8") if isinstance(s, str) else bytes(s)
        return Expr(self, lib().kq_expr_literal_utf8(b, len(b)))

    def lit_null(self, t): return Expr(self, lib().kq_expr_literal_null(t))
    def binary(self, op, l, r): return Expr(self, lib().kq_expr_binary(OPS[op], l.h, r.h), (l, r))
    def cast(self, e, t): return Expr(self, lib().kq_expr_cast(e.h, t), (e,))

    def explain_filter_project(self, pred, exprs, types, nullable=None, compile=True) -> str:
        """CUDA source of the query-specific part of the kernel for this shape (kq_explain_filter_project)."""
        n = len(types)
        t = (C.c_int * n)(*types)
        nl = (C.c_int * n)(*(nullable or [0] * n))
        buf = C.create_string_buffer(1 << 18)
        st = lib().kq_explain_filter_project(pred.h if pred is not None else None, _expr_array(exprs), len(exprs), n, t, nl,
                                             int(bool(compile)), buf, len(buf))
        text = buf.value.decode("utf-8", "replace")
        if st != 0:
            raise KqError(st, text)
        return text


def _explain_hashagg(group_exprs, aggs, types, nullable=None, pred=None, compile=True) -> str:
    n = len(types)
    t = (C.c_int * n)(*types)
    nl = (C.c_int * n)(*(nullable or [0] * n))
    kinds = (C.c_int * max(len(aggs), 1))(*[AGGS[k] for k, _ in aggs])
    buf = C.create_string_buffer(1 << 18)
    st = lib().kq_explain_hashagg(pred.h if pred is not None else None, _expr_array(group_exprs), len(group_exprs), kinds,
                                  _expr_array([e for _, e in aggs]), len(aggs), n, t, nl, int(bool(compile)), buf, len(buf))
    text = buf.value.decode("utf-8", "replace")
    if st != 0:
        raise KqError(st, text)
    return text


Exprs.explain_hashagg = staticmethod(_explain_hashagg)


class Engine(Exprs):
    """The operator vocabulary bound to one Context (same method names as oracle/oracle.py)."""

    OPS, AGGS = OPS, AGGS

    def __init__(self, ctx: Context):
        self.ctx = ctx
        outer = self

        class _RB:
            @staticmethod
            def from_arrow(arrays, n_rows=-1):
                cols = [Column.from_arrow(outer.ctx, a) for a in arrays]
                return RecordBatch.from_columns(outer.ctx, cols, n_rows)

            @staticmethod
            def from_columns(cols, n_rows=-1):
                return RecordBatch.from_columns(outer.ctx, cols, n_rows)

        self.RecordBatch = _RB
        self.KqError = KqError
        self.type_of = _type_of

    # operators
    def project(self, exprs, batch: RecordBatch) -> RecordBatch:
        out = C.c_void_p()
        self.ctx.check(lib().kq_project(self.ctx.h, _expr_array(exprs), len(exprs), batch.h, C.byref(out)))
        return RecordBatch(self.ctx, out)

    def filter(self, pred, batch: RecordBatch, want_selection=False):
        out, sel = C.c_void_p(), C.c_void_p()
        self.ctx.check(lib().kq_filter(self.ctx.h, pred.h, batch.h, C.byref(out), C.byref(sel) if want_selection else None))
        if want_selection:
            return RecordBatch(self.ctx, out), Column(self.ctx, sel)
        return RecordBatch(self.ctx, out)

    def filter_project(self, pred, exprs, batch: RecordBatch) -> RecordBatch:
        out = C.c_void_p()
        self.ctx.check(lib().kq_filter_project(self.ctx.h, pred.h, _expr_array(exprs), len(exprs), batch.h, C.byref(out)))
        return RecordBatch(self.ctx, out)

    def filter_project_host(self, pred, exprs, cols, n, outs) -> int:
        """kq_filter_project_host: host Arrow buffers in, host result buffers out, streamed in chunks with
        H2D / kernel / D2H overlapped. cols = [(kq_type, validity_address or None, data_address)];
        outs = [(data_address, validity_address or None)] with room for n rows each. Returns the row count."""
        nc, no = len(cols), len(outs)
        types = (C.c_int * max(nc, 1))(*[c[0] for c in cols])
        val = (C.c_void_p * max(nc, 1))(*[c[1] for c in cols])
        dat = (C.c_void_p * max(nc, 1))(*[c[2] for c in cols])
        od = (C.c_void_p * max(no, 1))(*[o[0] for o in outs])
        ov = (C.c_void_p * max(no, 1))(*[o[1] for o in outs])
        rows = C.c_int64(0)
        has_v = any(c[1] for c in cols)
        self.ctx.check(lib().kq_filter_project_host(self.ctx.h, pred.h, _expr_array(exprs), len(exprs), nc, types,
                                                    val if has_v else None, dat, n, od, ov if any(o[1] for o in outs) else None,
                                                    C.byref(rows)))
        return rows.value

    def HashAggregate(self, group_exprs, aggs, pred=None, expected_groups=0):
        return HashAggregate(self, group_exprs, aggs, pred, expected_groups)

    @staticmethod
    def csv_header(text: bytes, has_headers=True):
        """CsvDataSource.schema() (Main.kt:328-356): (column names, detected delimiter); every column is Utf8."""
        names = C.create_string_buffer(1 << 16)
        n, d = C.c_int(), C.create_string_buffer(2)
        st = lib().kq_csv_header(text, len(text), int(bool(has_headers)), names, len(names), C.byref(n), d)
        if st != 0:
            raise KqError(st, "kq_csv_header")
        return names.value.decode("utf-8").split("\n")[:n.value], d.raw[:1].decode()

    def csv_scan(self, text: bytes, has_headers=True, projection=None) -> RecordBatch:
        """CsvDataSource.scan(projection) (Main.kt:304-326): the text of a CSV file -> one batch of Utf8 columns on the
        device. `projection`: column NAMES in output order (Schema.select, Main.kt:47-52); None/[] = all columns."""
        idx = []
        if projection:
            names, _ = self.csv_header(text, has_headers)
            for p in projection:
                if p not in names:
                    raise KqError(3, f"Field {p} not found")        # KQ_ERR_ILLEGAL_ARGUMENT = IllegalArgumentException, Main.kt:49
                idx.append(names.index(p))
        arr = (C.c_int * max(len(idx), 1))(*idx)
        out = C.c_void_p()
        self.ctx.check(lib().kq_csv_scan(self.ctx.h, text, len(text), int(bool(has_headers)), arr if idx else None, len(idx), C.byref(out)))
        return RecordBatch(self.ctx, out)

    def csv_batches(self, text, has_headers=True, projection=None, piece_bytes=0, nbytes=None, columns=None):
        """CsvDataSource.scan(projection) as the reference returns it: a Sequence<RecordBatch> (ReaderIterator,
        Main.kt:239-249). A generator of device batches, one per piece of `piece_bytes` of text (0 = 64 MiB), each cut at a
        record boundary; the copy of the next piece runs under the scan of the current one. `text`: bytes, or an address
        (int: pinned/pageable host memory or a device pointer) with `nbytes`; `projection`: column names (bytes input only),
        `columns`: file column indices."""
        idx = list(columns or [])
        if isinstance(text, (bytes, bytearray)):
            nbytes = len(text)
            if projection:
                names, _ = self.csv_header(bytes(text), has_headers)
                for p in projection:
                    if p not in names:
                        raise KqError(3, f"Field {p} not found")        # Main.kt:49
                    idx.append(names.index(p))
            src = C.c_char_p(bytes(text))           # kept alive by this frame until the reader is closed
        else:
            src = C.cast(C.c_void_p(int(text)), C.c_char_p)
        arr = (C.c_int * max(len(idx), 1))(*idx)
        rd = C.c_void_p()
        self.ctx.check(lib().kq_csv_reader_open(self.ctx.h, src, nbytes, int(bool(has_headers)), arr if idx else None, len(idx),
                                                int(piece_bytes), C.byref(rd)))
        try:
            while True:
                out = C.c_void_p()
                self.ctx.check(lib().kq_csv_reader_next(rd, C.byref(out)))
                if not out.value:
                    return
                yield RecordBatch(self.ctx, out)
        finally:
            lib().kq_csv_reader_close(rd)

    def csv_scan_ptr(self, ptr: int, nbytes: int, has_headers=True, columns=None) -> RecordBatch:
        """kq_csv_scan on a raw buffer: pinned/pageable host memory or a device pointer (text already in HBM);
        `columns`: file column indices in output order."""
        idx = list(columns or [])
        arr = (C.c_int * max(len(idx), 1))(*idx)
        out = C.c_void_p()
        self.ctx.check(lib().kq_csv_scan(self.ctx.h, C.cast(C.c_void_p(ptr), C.c_char_p), nbytes, int(bool(has_headers)),
                                         arr if idx else None, len(idx), C.byref(out)))
        return RecordBatch(self.ctx, out)

    def generate(self, specs, seed, row_begin, row_end) -> RecordBatch:
        arr, keep = make_specs(specs)
        out = C.c_void_p()
        self.ctx.check(lib().kq_generate(self.ctx.h, arr, len(specs), seed, row_begin, row_end, C.byref(out)))
        return RecordBatch(self.ctx, out)
