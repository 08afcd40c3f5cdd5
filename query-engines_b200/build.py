"""Build libkqgpu.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python query-engines_b200/build.py [--force] [--verbose]

The shared object is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
OUT = os.path.join(HERE, "libkqgpu.so")
OBJDIR = os.path.join(HERE, "build")

SOURCES = ["kq_core.cu", "kq_compile.cu", "kq_ops.cu", "kq_hashagg.cu", "kq_comm.cu"]
HEADERS = ["kq_internal.h", "kq_vm.cuh", "kq_scan.cuh", "kq_compile.h"]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",                      # never contract a*b+c into FMA (rule E4)
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
    "-I", INCLUDE,
]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS if os.path.exists(os.path.join(CSRC, f))]
    deps += [os.path.join(INCLUDE, "kqgpu.h"), os.path.join(INCLUDE, "kq_gen.h"), os.path.abspath(__file__)]
    return _newest(deps) > os.path.getmtime(OUT)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(src):
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(os.path.join(OBJDIR, src + ".log"), "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(compile_one, srcs))
    link = [NVCC, "-shared", "-o", OUT, *objs, "-cudart", "static", "-ldl", "-lpthread"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
