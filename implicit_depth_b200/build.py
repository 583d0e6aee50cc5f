"""Build the C-ABI library ``csrc/liblidf_query.so`` in-tree with nvcc for sm_100a.

The library has no torch / Python dependency: plain ``extern "C"`` entry points declared in
``include/lidf_query.h``.  nvcc cross-compiles without a GPU, so this also is the CPU-side build check.
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(REPO, "include")
# LIDF_QUERY_LIB points the loader at another build of the same sources (A/B timing experiments); it is never rebuilt
LIB_OVERRIDE = os.environ.get("LIDF_QUERY_LIB")
LIB_PATH = LIB_OVERRIDE or os.path.join(CSRC, "liblidf_query.so")
SOURCES = ["lidf_query.cu"]
HEADERS = ["lidf_common.cuh", "lidf_prep.cuh", "lidf_simt.cuh", "lidf_tc.cuh", "lidf_bwd.cuh", "lidf_aabb.cuh",
           "lidf_pointnet.cuh", os.path.join(INCLUDE, "lidf_query.h"), os.path.join(INCLUDE, "lidf_aabb.h"),
           os.path.join(INCLUDE, "lidf_pointnet.h")]
# no --use_fast_math: the path needs the accurate sincosf / expf / IEEE division
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC"]

def find_nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if missing or older than its sources.  Returns the library path."""
    if LIB_OVERRIDE or (not force and not is_stale()):
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build liblidf_query.so (no CPU fallback exists for this path)")
    cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-I", CSRC, "-o", LIB_PATH + ".tmp"] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas"); cmd.insert(2, "-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
