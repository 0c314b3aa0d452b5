"""Builds libbspy_cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m bspy_b200._cuda.build [--force] [--verbose]

The shared library is a plain C-ABI object (include/bspy_cuda.h): no torch, no pybind.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbspy_cuda.so")
STAMP = LIB + ".stamp"
SOURCES = ["core.cu", "scattered.cu", "cells.cu", "curve.cu", "grid.cu", "many.cu", "block.cu", "probe.cu"]
HEADERS = ["common.cuh", "curve.cuh", "scattered.cuh", os.path.join("..", "..", "..", "include", "bspy_cuda.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libbspy_cuda.so")
    return exe


def _digest(paths):
    import hashlib
    h = hashlib.sha256(" ".join(ARCH + FLAGS).encode())
    for path in paths:
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _header_paths():
    return [os.path.join(CSRC, f) for f in HEADERS]


def _library_digest():
    return _digest([os.path.join(CSRC, f) for f in SOURCES] + _header_paths())


def stale():
    """True when the library is missing or was built from other sources (content hash, not mtimes: the tree is
    copied to the GPU box and file times need not survive the trip)."""
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _library_digest()


def build(force=False, verbose=False):
    """Compile and link when the library is missing or older than its sources.  Safe under torchrun (one process
    per GPU importing the package at once): an inter-process file lock serialises the builders, objects and the
    library are written under temporary names and renamed into place, and the losers of the race find a fresh
    library when they get the lock."""
    if not force and not stale():
        return LIB
    import fcntl
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    with open(os.path.join(objdir, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not stale():
                return LIB
            return _build_locked(objdir, verbose, force)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(objdir, verbose, force=False):
    tag = f".{os.getpid()}.tmp"
    procs, objs, log = [], [], []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        want = _digest([os.path.join(CSRC, src)] + _header_paths())
        stamp = obj + ".stamp"
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == want:
            continue                                        # this object is current: only changed sources recompile
        cmd = [nvcc(), *ARCH, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj + tag]
        procs.append((src, obj, want, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, obj, want, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        with open(os.path.join(objdir, src.replace(".cu", ".ptxas.log")), "w") as f:
            f.write(out)
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
        os.replace(obj + tag, obj)
        with open(obj + ".stamp", "w") as f:
            f.write(want)
    if verbose:
        print("\n".join(log))
    subprocess.check_call([nvcc(), *ARCH, "-shared", "-o", LIB + tag, *objs, "-lcudart"])
    os.replace(LIB + tag, LIB)
    with open(STAMP, "w") as f:
        f.write(_library_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
