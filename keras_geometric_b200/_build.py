"""Build libkgb200.so (hand-written sm_100a CUDA behind the C ABI in include/kgb200.h).

nvcc cross-compiles without a GPU; the .so is kept in-tree next to this file so it travels with
the repository snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(PKG, "libkgb200.so")
OBJ_DIR = os.path.join(CSRC, "build")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-I", INCLUDE, "-I", CSRC,
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: libkgb200.so cannot be built on this machine")
    return cand


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    files = _sources() + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")]
    files.append(os.path.join(INCLUDE, "kgb200.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode())  # location-independent: the tree is copied to the GPU box
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).replace(ROOT, "<root>").encode())
    return h.hexdigest()


def _file_digest(src: str, cmd) -> str:
    import re
    h = hashlib.sha256()
    deps, todo = [], [src]
    while todo:  # transitive closure of the quoted includes that live in csrc/ or include/
        f = todo.pop()
        if f in deps:
            continue
        deps.append(f)
        with open(f) as fh:
            for inc in re.findall(r'#include\s+"([^"]+)"', fh.read()):
                for base in (CSRC, INCLUDE):
                    cand = os.path.join(base, inc)
                    if os.path.exists(cand):
                        todo.append(cand)
    for f in sorted(deps):
        with open(f, "rb") as fh:
            h.update(fh.read())
    import sysconfig
    norm = " ".join(cmd).replace(ROOT, "<root>")
    for base in {sysconfig.get_paths().get("purelib"), sysconfig.get_paths().get("platlib")}:
        if base:
            norm = norm.replace(base, "<site>")
    h.update(norm.encode())
    return h.hexdigest()


def is_fresh() -> bool:
    stamp = LIB + ".sha256"
    if not (os.path.exists(LIB) and os.path.exists(stamp)):
        return False
    with open(stamp) as fh:
        return fh.read().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a and link libkgb200.so.  Returns the library path."""
    if not force and is_fresh():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    # one builder at a time (several ranks may import the package at once)
    import fcntl
    lock = open(os.path.join(OBJ_DIR, ".lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and is_fresh():
            return LIB
        return _build_locked(nvcc, force, verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(nvcc: str, force: bool, verbose: bool) -> str:

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        if os.path.exists(obj) and os.path.exists(obj + ".sha"):  # per-object cache
            with open(obj + ".sha") as fh:
                if fh.read() == _file_digest(src, cmd):
                    return obj
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        with open(obj + ".sha", "w") as fh:
            fh.write(_file_digest(src, cmd))
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(_sources()))) as ex:
        objs = list(ex.map(compile_one, _sources()))
    tmp = LIB + ".tmp"
    cmd = [nvcc, "-shared", "-o", tmp] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)  # atomic: a concurrent loader never sees a half-written library
    with open(LIB + ".sha256", "w") as fh:
        fh.write(_digest())
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
