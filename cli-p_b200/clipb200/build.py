"""Build libclipb200.so in-tree with plain nvcc for sm_100a (cross-compiles without a GPU).

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libclipb200.so")
STAMP = os.path.join(HERE, ".libclipb200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../../include/clipb200.h"]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            with open(p, "rb") as fh:
                h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force: bool = False, verbose: bool = False, experiments: bool = False) -> str:
    """experiments=True builds libclipb200_exp.so with -DCLIPB200_EXPERIMENTS: the same library plus the
    result-corrupting perf probes (gemm_debug / skip knobs) that profiles/gemm_decompose.py uses.  Load
    it with CLIPB200_LIB=<path>; the product library never contains them."""
    if experiments:
        return _build_variant(os.path.join(HERE, "libclipb200_exp.so"), os.path.join(HERE, "build_exp"),
                              ["-DCLIPB200_EXPERIMENTS"], verbose)
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB
    _build_variant(LIB, os.path.join(HERE, "build"), [], verbose)
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


def _build_variant(lib: str, objdir: str, extra, verbose: bool) -> str:
    objs = []
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc_path(), *NVCC_FLAGS, *extra, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {src}\n")
    if failed:
        raise RuntimeError("libclipb200 build failed")
    link = [nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("libclipb200 link failed")
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, experiments="--experiments" in sys.argv))
