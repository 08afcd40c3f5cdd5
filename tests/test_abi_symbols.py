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
