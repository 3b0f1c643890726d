"""Builds the C-ABI CUDA libraries in-tree with nvcc for sm_100a.

    mmmm_b200/libvex.so            the product: every kernel the visual-expert path launches (csrc/*.cu)
    mmmm_b200/libvex_baselines.so  the superseded attention kernels kept for A/B runs (csrc/baselines/*.cu); only
                                   tests/ and tools/ load it

Neither library depends on torch (plain CUDA runtime, statically linked), so they compile in seconds and load through
ctypes.  ``python -m mmmm_b200.build`` or ``build(force=True)``.

Concurrency: under torchrun every rank may find the library stale at the same time.  The build takes an inter-process
file lock, compiles objects and links into temporary names and publishes them with ``os.replace`` (atomic), so a
rank can never ``dlopen`` a half-written file; the ranks that waited find the stamp fresh and skip the build.
"""
from __future__ import annotations

import contextlib
import fcntl
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BASE = os.path.join(CSRC, "baselines")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libvex.so")
LIB_BASELINES = os.path.join(HERE, "libvex_baselines.so")
STAMP = os.path.join(OBJ, "stamp.txt")
STAMP_BASELINES = os.path.join(OBJ, "stamp_baselines.txt")
LOCK = os.path.join(OBJ, ".lock")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-Xptxas", "-v", "-Xcudafe", "--diag_suppress=177"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def baseline_sources():
    return sorted(os.path.join("baselines", f) for f in os.listdir(BASE) if f.endswith(".cu")) + ["tmap.cu"]


def _digest(roots) -> str:
    h = hashlib.sha256()
    for root in roots:
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(ARCH_FLAGS + CFLAGS).encode())
    return h.hexdigest()


_INC = os.path.join(os.path.dirname(HERE), "include")


def _fresh(lib: str, stamp: str, roots) -> bool:
    if not (os.path.isfile(lib) and os.path.isfile(stamp)):
        return False
    with open(stamp) as f:
        return f.read().strip() == _digest(roots)


def is_fresh() -> bool:
    return _fresh(LIB, STAMP, (CSRC, _INC))


def baselines_fresh() -> bool:
    return _fresh(LIB_BASELINES, STAMP_BASELINES, (CSRC, BASE, _INC))


@contextlib.contextmanager
def _locked():
    os.makedirs(OBJ, exist_ok=True)
    with open(LOCK, "w") as fh:
        fcntl.flock(fh, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(fh, fcntl.LOCK_UN)


def _compile(src: str) -> str:
    stem = src[:-3].replace(os.sep, "_")
    obj = os.path.join(OBJ, stem + ".o")
    tmp = f"{obj}.{os.getpid()}.tmp"
    cmd = [NVCC, *ARCH_FLAGS, *CFLAGS, "-c", os.path.join(CSRC, src), "-o", tmp]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(OBJ, stem + ".log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, obj)
    return obj


def _build_one(lib: str, stamp: str, srcs, roots, verbose: bool) -> str:
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(_compile, srcs))
    tmp = f"{lib}.{os.getpid()}.tmp"
    cmd = [NVCC, *ARCH_FLAGS, "-shared", "-Xcompiler", "-fPIC", "-o", tmp, *objs, "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, lib)  # atomic: a concurrent dlopen sees the old or the new file, never a partial one
    with open(stamp + ".tmp", "w") as f:
        f.write(_digest(roots))
    os.replace(stamp + ".tmp", stamp)
    if verbose:
        for s in srcs:
            with open(os.path.join(OBJ, s[:-3].replace(os.sep, "_") + ".log")) as f:
                print(f.read())
    return lib


def _require_nvcc():
    if not os.path.isfile(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}; libvex.so cannot be built (no CPU fallback exists)")


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and is_fresh():
        return LIB
    _require_nvcc()
    with _locked():
        if not force and is_fresh():  # another process built it while this one waited for the lock
            return LIB
        return _build_one(LIB, STAMP, sources(), (CSRC, _INC), verbose)


def build_baselines(force: bool = False, verbose: bool = False) -> str:
    if not force and baselines_fresh():
        return LIB_BASELINES
    _require_nvcc()
    with _locked():
        if not force and baselines_fresh():
            return LIB_BASELINES
        return _build_one(LIB_BASELINES, STAMP_BASELINES, baseline_sources(), (CSRC, BASE, _INC), verbose)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_baselines(force="--force" in sys.argv, verbose="-v" in sys.argv))
