"""Builds libgridap_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

    python gridap.jl_b200/build.py [--force]

The library has no dependency on torch / Python: it is what a Julia `ccall` (or any FFI) binds.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(OUT_DIR, "libgridap_b200.so")
SOURCES = ["api.cu", "symbolic.cu", "element_kernels.cu", "q1hex_gather.cu", "affine_gather.cu", "q1hex_rhs.cu", "vector_kernels.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-O2", "-Xcompiler", "-pthread", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "gridap_b200.h"))
    return hdrs


def _stale(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + _deps()):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        r = subprocess.run([NVCC] + FLAGS + ["-c", s, "-o", o], capture_output=True, text=True)
        log = os.path.join(OUT_DIR, os.path.basename(o) + ".ptxas.log")
        with open(log, "w") as f:   # register / spill report of every kernel (tracked); compile times dropped: they differ per build
            f.write("".join(l for l in (r.stdout + r.stderr).splitlines(True) if "Compile time" not in l))
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (s, r.stderr[-4000:]))
        return r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        for out in ex.map(compile_one, jobs):
            if verbose:
                print(out)
    if force or jobs or _stale(LIB, objs):
        r = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart", "-lpthread"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
