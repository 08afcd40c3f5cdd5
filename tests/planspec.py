"""Engine-neutral expression specs for the parity tests.

A spec is a nested tuple:
    ("col", i) | ("lit", type_name, value) | ("bin", OP, l, r) | ("cast", e, type_name)
build(E, spec) turns it into the expression objects of engine module E (oracle.oracle or kqgpu),
np_ref.evaluate(spec, columns) evaluates it independently with numpy.
"""

TYPES = {"f64": 1, "utf8": 2, "i64": 3, "bool": 4, "date32": 5, "i32": 6}


def build(E, spec):
    k = spec[0]
    if k == "col":
        return E.col(spec[1])
    if k == "lit":
        t, v = spec[1], spec[2]
        if v is None:
            return E.lit_null(TYPES[t])
        return {"f64": E.lit_f64, "i64": E.lit_i64, "bool": E.lit_bool, "date32": E.lit_date32,
                "utf8": E.lit_utf8}[t](v)
    if k == "bin":
        return E.binary(spec[1], build(E, spec[2]), build(E, spec[3]))
    if k == "cast":
        return E.cast(build(E, spec[1]), TYPES[spec[2]])
    raise ValueError(spec)


def col(i): return ("col", i)
def lit(t, v): return ("lit", t, v)
def b(op, l, r): return ("bin", op, l, r)


def sort_rows(arrays, nkeys):
    """Rows of a result (list of pyarrow arrays) as a list of tuples sorted by the first nkeys
    columns (None sorts first) — HashAggregate output order is unspecified (rule R10/E8)."""
    cols = [a.to_pylist() for a in arrays]
    rows = list(zip(*cols)) if cols else []
    def keyf(r):
        return tuple((0, "") if v is None else (1, v) for v in r[:nkeys])
    return sorted(rows, key=keyf)
