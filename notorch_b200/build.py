"""Build ``libnotorch_b200.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m notorch_b200.build [--force]

The shared library is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libnotorch_b200.so")

SOURCES = ["api.cu", "index_kernels.cu", "rowwise_kernels.cu", "gemm_simt.cu", "gemm_pair.cu", "wgrad_tc.cu", "wgrad_pair.cu", "embed_kernels.cu", "embed_fused.cu", "readout_kernels.cu", "pooled_backward.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libnotorch_b200.so cannot be built")


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join(INCLUDE, "notorch_b200.h"), os.path.join(INCLUDE, "notorch_b200_debug.h")]
    for f in files:
        path = f if os.path.isabs(f) else os.path.join(CSRC, f)
        with open(path, "rb") as fh:
            h.update(os.path.basename(f).encode())  # never the absolute path: the tree is copied to another location on the GPU box
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(" ".join(SOURCES).encode())  # a file that exists in csrc/ but is not linked yet must not read as "fresh"
    return h.hexdigest()


def is_fresh() -> bool:
    stamp = os.path.join(BUILD, "digest")
    return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and is_fresh():
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    # one builder at a time (torchrun starts N ranks that all import the package): the others wait, then find the library fresh
    import fcntl

    with open(os.path.join(BUILD, "lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_fresh():
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()

    def compile_one(src: str) -> tuple[str, str]:
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    with open(os.path.join(BUILD, "ptxas.log"), "w") as fh:
        for (_, log), src in zip(results, SOURCES):
            fh.write(f"==== {src}\n{log}\n")
            if verbose:
                print(f"==== {src}\n{log}")
    objs = [o for o, _ in results]
    tmp = LIB + ".tmp"  # link beside the target, then rename: a concurrent dlopen never sees a half-written library
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    with open(os.path.join(BUILD, "digest"), "w") as fh:
        fh.write(_digest())
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
