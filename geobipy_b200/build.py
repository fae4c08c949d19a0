"""In-tree build of the CUDA extension (nvcc, sm_100a only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "gbp_api.cu")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("gbp_api.cu", "gbp_chain.cuh", "gbp_fdem.cuh", "gbp_fdem_f2.cuh", "gbp_tdem.cuh", "gbp_tdem_f2.cuh", "gbp_math.cuh", "gbp_tables.h",
                                                     "gbp_tdem_tables.h")] + [
    os.path.join(ROOT, "include", f) for f in ("geobipy_b200.h", "gbp_filter_tables.h")]
OUT = os.path.join(HERE, "libgeobipy_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    """Compile geobipy_b200/libgeobipy_b200.so.  Cross-compiles without a GPU."""
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    subprocess.check_call(cmd)
    return OUT
