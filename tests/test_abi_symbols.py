"""CPU checks of the drop-in boundary: libkqgpu.so loads without a GPU, exports every symbol that
include/kqgpu.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kqgpu.h")


@pytest.fixture(scope="module")
def libpath():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "query-engines_b200"))
    import build
    return build.build()


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"KQ_API\s+[\w\s\*]+?\b(kq_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    assert len(names) >= 50
    for must in ("kq_ctx_create", "kq_column_upload", "kq_expr_evaluate", "kq_project", "kq_filter",
                 "kq_filter_project", "kq_hashagg_update", "kq_hashagg_merge_allreduce",
                 "kq_hashagg_repartition_alltoall", "kq_generate"):
        assert must in names


def test_library_exports_every_declared_symbol(libpath):
    out = subprocess.check_output(["nm", "-D", "--defined-only", libpath], text=True)
    exported = set(re.findall(r"\sT\s+(kq_\w+)", out))
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, f"declared in kqgpu.h but not exported: {missing}"


def test_binding_table_matches_header(libpath):
    import kqgpu
    assert sorted(kqgpu.SYMBOLS) == declared_symbols()
    kqgpu.lib()   # resolves every symbol through ctypes


def test_library_is_sm100a_native(libpath):
    out = subprocess.run(["cuobjdump", "-lelf", libpath], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="CPU-only check")
def test_no_cpu_fallback(libpath):
    import kqgpu
    assert kqgpu.device_count() == 0
    with pytest.raises(kqgpu.KqError) as e:
        kqgpu.Context(0)
    assert e.value.code == 10 and e.value.exception_class == "NoCudaDevice"
    # expression handles are pure host objects and work without a device
    L = kqgpu.lib()
    ex = L.kq_expr_binary(5, L.kq_expr_column(0), L.kq_expr_literal_i64(7))
    assert ex
    L.kq_expr_free(ex)


def test_kotlin_ffm_descriptors_match_the_binding_table(libpath):
    """INTEGRATION.md's Kotlin shim cannot be compiled here (no JDK), but its Panama FunctionDescriptors can be read: every
    `fn("kq_...", RET, ARGS...)` must name an exported symbol and agree with the ctypes table (itself checked against the
    header) in arity and in the class of every argument — 32-bit int, 64-bit int, double, float, address."""
    import ctypes as C
    import kqgpu
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    found = re.findall(r'fn\("(kq_\w+)",\s*([^)]*)\)', text)
    assert len(found) >= 25

    def cls(t):
        if t is None:
            return "void"
        if t in (C.c_int, C.c_int32, C.c_uint32):
            return "JAVA_INT"
        if t in (C.c_int64, C.c_uint64, C.c_size_t):
            return "JAVA_LONG"
        if t is C.c_double:
            return "JAVA_DOUBLE"
        if t is C.c_float:
            return "JAVA_FLOAT"
        return "ADDRESS"            # c_void_p, c_char_p, POINTER(...)

    for name, sig in found:
        assert name in kqgpu.SYMBOLS, f"INTEGRATION.md binds {name}, which include/kqgpu.h does not declare"
        parts = [p.strip() for p in sig.split(",") if p.strip()]
        ret, args = parts[0], parts[1:]
        res, argtypes = kqgpu.SYMBOLS[name]
        want = [cls(res)] + [cls(a) for a in argtypes]
        got = ["void" if ret == "null" else ret] + args
        assert got == want, f"{name}: Kotlin descriptor {got} != C signature {want}"


def test_binding_table_matches_the_header_signatures():
    """The ctypes table against the C declarations themselves: return class and the class of every parameter."""
    import ctypes as C
    import kqgpu
    text = re.sub(r"/\*.*?\*/", " ", open(HEADER).read(), flags=re.S)
    decls = re.findall(r"KQ_API\s+([\w\s\*]+?)\b(kq_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    assert len(decls) == len(declared_symbols())

    def c_class(decl):
        decl = decl.strip()
        if "*" in decl or "[" in decl:          # an array parameter is a pointer
            return "ptr"
        base = decl.replace("const", " ").split()
        base = [w for w in base if w not in ("unsigned", "signed")]
        if decl == "void":
            return "void"
        ty = base[0]
        return {"int": "i32", "int32_t": "i32", "uint32_t": "i32", "int64_t": "i64", "uint64_t": "i64", "size_t": "i64",
                "double": "f64", "float": "f32"}[ty]

    def py_class(t):
        if t is None:
            return "void"
        if t in (C.c_int, C.c_int32, C.c_uint32):
            return "i32"
        if t in (C.c_int64, C.c_uint64, C.c_size_t):
            return "i64"
        if t is C.c_double:
            return "f64"
        if t is C.c_float:
            return "f32"
        return "ptr"

    for ret, name, params in decls:
        params = [p.strip() for p in params.replace("\n", " ").split(",")]
        params = [] if params == ["void"] else params
        want = [c_class(ret)] + [c_class(p) for p in params]
        res, argtypes = kqgpu.SYMBOLS[name]
        got = [py_class(res)] + [py_class(a) for a in argtypes]
        assert got == want, f"{name}: ctypes {got} != header {want}"
