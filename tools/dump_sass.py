"""Compile a BASELINE aggregate shape without a GPU and keep source + cubin for cuobjdump (KQ_JIT_DUMP).
   python tools/dump_sass.py cfg3|cfg4|cfg5 /tmp/out_prefix [expected_groups]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200"))
which, prefix = sys.argv[1], sys.argv[2]
os.environ["KQ_JIT_DUMP"] = prefix
import build; build.build()
import kqgpu
E = kqgpu.Exprs()
F64, UTF8, I64, BOOL, D32 = 1, 2, 3, 4, 5
v = E.col(1)
four = [("SUM", v), ("MIN", v), ("MAX", v), ("COUNT", v)]
if which == "cfg3":
    src = E.explain_hashagg([E.col(0)], four, [UTF8, F64])
elif which == "cfg4":
    src = E.explain_hashagg([E.col(0)], four, [I64, F64])
else:
    one = E.lit_f64(1.0)
    dp = E.binary("MUL", E.col(4), E.binary("SUB", one, E.col(5)))
    ch = E.binary("MUL", dp, E.binary("ADD", one, E.col(6)))
    src = E.explain_hashagg([E.col(1), E.col(2)], [("SUM", E.col(3)), ("SUM", E.col(4)), ("SUM", dp), ("SUM", ch), ("COUNT", E.lit_i64(1))],
                            [D32, UTF8, UTF8, F64, F64, F64, F64], pred=E.binary("LE", E.col(0), E.lit_date32(10471)))
print(src[:600])
