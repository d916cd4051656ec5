"""Build the CUDA library in-tree: compas_tno_b200/libtno_b200.so (sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtno_b200.so")
SOURCES = ["tno_capi.cu", "tno_interleaved.cu", "tno_onchip.cu", "tno_util.cu", "tno_sqp.cu", "tno_preprocess.cu", "tno_symbolic.cpp"]
HEADERS = ["tno_plan.hpp", "tno_symbolic.hpp", os.path.join("..", "..", "include", "tno_b200.h")]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


STAMP = os.path.join(HERE, "libtno_b200.srchash")


def source_hash():
    """sha1 over the sources and headers the library is built from (content, not mtime: the tree is copied to the GPU
    box, which does not keep time stamps in order)."""
    import hashlib
    h = hashlib.sha1()
    for name in SOURCES + HEADERS:
        path = os.path.join(CSRC, name)
        if os.path.exists(path):
            h.update(name.encode())
            with open(path, "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB):
        return True
    try:
        with open(STAMP) as fh:
            return fh.read().strip() != source_hash()
    except OSError:
        return True


def build(force=False, verbose=False):
    """Compile every translation unit (in parallel, one nvcc process each) and link the shared library."""
    if not force and not needs_build():
        return LIB
    import fcntl
    from concurrent.futures import ThreadPoolExecutor
    # one build at a time: the ranks of a torchrun job and the workers of pytest-xdist share this directory
    lock = open(os.path.join(HERE, ".build.lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and not needs_build():
            return LIB
        return _build_locked(verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(verbose):
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(HERE, "..", "build", "obj")
    os.makedirs(objdir, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    flags = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC,-O3"]
    if verbose:
        flags += ["-Xptxas", "-v"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        r = subprocess.run([_nvcc()] + flags + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=len(srcs)) as pool:
        results = list(pool.map(compile_one, srcs))
    for src, _, r in results:
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on %s" % src)
        if verbose:
            sys.stderr.write(r.stderr)
    tmp = LIB + ".tmp.%d" % os.getpid()
    r = subprocess.run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + [o for _, o, _ in results],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc link failed")
    os.replace(tmp, LIB)                 # atomic: a process that is loading the old library keeps its mapping
    with open(STAMP, "w") as fh:
        fh.write(source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
