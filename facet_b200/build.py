"""Build libfacet_b200.so in-tree with nvcc for sm_100a (no torch headers involved).

    python -m facet_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfacet_b200.so")
STAMP = os.path.join(HERE, ".libfacet_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    paths = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    paths.append(os.path.join(os.path.dirname(HERE), "include", "facet_b200.h"))
    for p in paths:
        if os.path.isfile(p):
            h.update(os.path.basename(p).encode())      # not the absolute path: the tree is relocated on the GPU box
            with open(p, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: facet_b200 needs the CUDA toolkit to build its kernels")


def _up_to_date(digest: str) -> bool:
    if os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            return f.read().strip() == digest
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    import fcntl
    digest = _digest()
    if not force and _up_to_date(digest):
        return LIB
    # several ranks may import the package at once: one builds, the others wait on the lock and re-check
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _up_to_date(digest):
                return LIB
            return _build_locked(digest, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(digest: str, verbose: bool) -> str:
    nvcc = nvcc_path()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
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
        raise RuntimeError("building libfacet_b200.so failed")
    tmp_lib = LIB + ".tmp"
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o", tmp_lib, *objs]
    subprocess.check_call(link)
    os.replace(tmp_lib, LIB)            # atomic: a concurrent loader never sees a half-written file
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
