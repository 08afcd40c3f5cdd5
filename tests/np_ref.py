"""Independent numpy restatement of the operator semantics (rules R1-R12, E1-E8 of SURVEY.md §8c).

Second opinion for the C++ oracle: vectorised, shares no code with it. A column is a pair
(values: np.ndarray, valid: np.ndarray[bool]); Utf8 values are an object array of bytes.
"""
import numpy as np
import pyarrow as pa


def from_arrow(arr):
    if isinstance(arr, pa.ChunkedArray):
        arr = arr.combine_chunks()
    valid = np.array([v is not None for v in arr.to_pylist()], dtype=bool) if arr.null_count else np.ones(len(arr), bool)
    t = arr.type
    if t == pa.string():
        vals = np.array([(v.encode("utf-8") if v is not None else b"") for v in arr.to_pylist()], dtype=object)
        return ("utf8", vals, valid)
    if t == pa.float64():
        return ("f64", np.nan_to_num(arr.fill_null(0.0).to_numpy(zero_copy_only=False), nan=np.nan), valid)
    if t == pa.int64():
        return ("i64", arr.fill_null(0).to_numpy(zero_copy_only=False).astype(np.int64), valid)
    if t == pa.bool_():
        return ("bool", arr.fill_null(False).to_numpy(zero_copy_only=False).astype(bool), valid)
    if t == pa.date32():
        return ("date32", arr.cast(pa.int32()).fill_null(0).to_numpy(zero_copy_only=False).astype(np.int64), valid)
    raise TypeError(t)


def to_pylist(col):
    t, v, ok = col
    out = []
    for x, k in zip(v.tolist(), ok.tolist()):
        if not k:
            out.append(None)
        elif t == "utf8":
            out.append(x.decode("utf-8"))
        elif t == "bool":
            out.append(bool(x))
        else:
            out.append(x)
    return out


def evaluate(spec, cols, n):
    k = spec[0]
    if k == "col":
        return cols[spec[1]]
    if k == "lit":
        t, v = spec[1], spec[2]
        ok = np.full(n, v is not None, dtype=bool)
        if t == "utf8":
            b = (v.encode("utf-8") if isinstance(v, str) else (v or b""))
            return (t, np.array([b] * n, dtype=object), ok)
        dt = {"f64": np.float64, "i64": np.int64, "bool": bool, "date32": np.int64}[t]
        return (t, np.full(n, v if v is not None else 0, dtype=dt), ok)
    if k == "cast":
        t, v, ok = evaluate(spec[1], cols, n)
        assert spec[2] == "f64"
        if t == "utf8":
            return ("f64", np.array([float(x) if o else 0.0 for x, o in zip(v, ok)], dtype=np.float64), ok)
        return ("f64", v.astype(np.float64), ok)
    if k == "bin":
        op = spec[1]
        ta, a, oa = evaluate(spec[2], cols, n)
        tb, b, ob = evaluate(spec[3], cols, n)
        assert ta == tb, "E2: operand types must match"
        if op in ("AND", "OR"):
            at, af = oa & a, oa & ~a
            bt, bf = ob & b, ob & ~b
            if op == "AND":
                false_, true_ = af | bf, at & bt
            else:
                true_, false_ = at | bt, af & bf
            return ("bool", true_, true_ | false_)
        ok = oa & ob
        if op in ("EQ", "NE", "LT", "LE", "GT", "GE"):
            if ta == "utf8":
                f = {"EQ": lambda x, y: x == y, "NE": lambda x, y: x != y, "LT": lambda x, y: x < y,
                     "LE": lambda x, y: x <= y, "GT": lambda x, y: x > y, "GE": lambda x, y: x >= y}[op]
                r = np.array([f(x, y) for x, y in zip(a, b)], dtype=bool)
            else:
                with np.errstate(invalid="ignore"):
                    r = {"EQ": np.equal, "NE": np.not_equal, "LT": np.less, "LE": np.less_equal,
                         "GT": np.greater, "GE": np.greater_equal}[op](a, b)
            return ("bool", r & ok, ok)
        with np.errstate(all="ignore"):
            if op == "ADD": r = a + b
            elif op == "SUB": r = a - b
            elif op == "MUL": r = a * b
            else:
                if ta == "i64":
                    if np.any(ok & (b == 0)):
                        raise ZeroDivisionError("/ by zero")
                    bb = np.where(b == 0, 1, b)
                    q = np.abs(a.astype(object)) // np.abs(bb.astype(object))   # truncating division
                    q = np.where((a < 0) != (bb < 0), -q, q)
                    r = np.array([((int(x) + 2**63) % 2**64) - 2**63 for x in q], dtype=np.int64)
                else:
                    r = a / b
        return (ta, r, ok)
    raise ValueError(spec)


def filter_rows(pred_col):
    _, v, ok = pred_col
    return np.nonzero(v & ok)[0]


def take(col, idx):
    t, v, ok = col
    return (t, v[idx], ok[idx])


def group_aggregate(key_cols, aggs):
    """aggs: list of (KIND, column). Returns dict key-tuple -> list of final values (None = null)."""
    n = len(key_cols[0][1]) if key_cols else (len(aggs[0][1][1]) if aggs else 0)
    keys = [to_pylist(c) for c in key_cols]
    out = {}
    def canon(x):
        if isinstance(x, float):
            if x != x: return ("nan",)
            return (np.float64(x).tobytes(),)     # +0.0 and -0.0 are different groups (R7)
        return x
    for r in range(n):
        kt = tuple(canon(k[r]) for k in keys)
        if kt not in out:
            out[kt] = {"key": tuple(k[r] for k in keys), "st": [None] * len(aggs), "cnt": [0] * len(aggs)}
        e = out[kt]
        for j, (kind, (t, v, ok)) in enumerate(aggs):
            if not ok[r]:
                continue
            x = v[r].item() if hasattr(v[r], "item") else v[r]
            e["cnt"][j] += 1
            cur = e["st"][j]
            if cur is None:
                e["st"][j] = x
            elif kind == "MAX":
                if x > cur: e["st"][j] = x
            elif kind == "MIN":
                if x < cur: e["st"][j] = x
            elif kind == "SUM":
                e["st"][j] = (cur + x) if t == "f64" else ((cur + x + 2**63) % 2**64) - 2**63
    res = {}
    for kt, e in out.items():
        res[e["key"] if not any(isinstance(x, float) and x != x for x in e["key"]) else kt] = [
            (e["cnt"][j] if kind == "COUNT" else e["st"][j]) for j, (kind, _) in enumerate(aggs)]
    return res
