"""Build libmsda_b200.so in-tree with nvcc for sm_100a.

    python -m gomatching_b200.build [--force] [--verbose]

The library is plain CUDA C++ behind a C ABI (include/msda_b200.h): no torch headers, no pybind, so the
same .so serves the Python host layer (ctypes), the cgo/JNI stubs of INTEGRATION.md and C callers.
The output lives next to this file (git-ignored, shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmsda_b200.so")
SOURCES = ["msda_forward.cu", "msda_forward_fast.cu", "msda_forward_staged.cu", "msda_forward_pipelined.cu", "msda_forward_paired.cu", "msda_backward.cu", "msda_api.cu", "proj_gemm.cu", "frame_batcher.cu", "add_layernorm.cu", "small_attention.cu", "decoder_glue.cu", "frame_resize.cu", "jpeg_decode.cu"]
HEADERS = ["msda_device.cuh", "msda_fast_common.cuh", "tma_common.cuh", "msda_launch.h", "jpeg_entropy.h", os.path.join("..", "..", "include", "msda_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    # no --use_fast_math: division, expf and the explicit __f*_rn chain must stay IEEE (index contract)
]


STAMP = os.path.join(HERE, "libmsda_b200.srchash")


def _source_hash() -> str:
    """Content hash of everything the library is built from (mtimes do not survive the copy to the GPU box)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in sorted(os.path.join(CSRC, s) for s in SOURCES + HEADERS):
        if os.path.exists(d):
            with open(d, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def _stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _source_hash()


def have_nvcc() -> bool:
    return bool(shutil.which("nvcc")) or os.path.exists("/usr/local/cuda/bin/nvcc")


def build(force: bool = False, verbose: bool = False, diag: bool = False) -> str:
    """No-op when the library is newer than every source.  Concurrent callers (one rank per GPU on a fresh checkout)
    are serialised by a file lock, and objects / the link output go to per-process temporaries before one atomic rename."""
    if not force and not _stale():
        return LIB
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():               # another process built it while this one waited
                return LIB
            return _build_locked(verbose, diag)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool, diag: bool = False) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libmsda_b200.so cannot be built (there is no CPU fallback)")
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(CSRC, s.replace(".cu", ".%d.o" % os.getpid()))
        cmd = [nvcc, *NVCC_FLAGS, *(["-DMSDA_DIAG"] if diag else []), "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, pr in procs:
        out, _ = pr.communicate()
        if verbose and out:
            print(out)
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (s, out))
    tmp = LIB + ".%d.tmp" % os.getpid()
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    os.replace(tmp, LIB)
    with open(STAMP, "w") as f:
        f.write((_source_hash() if not diag else "diagnostic build: always stale") + "\n")
    for o in objs:
        try:
            os.remove(o)
        except OSError:
            pass
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv or "--diag" in sys.argv, verbose="--verbose" in sys.argv, diag="--diag" in sys.argv)
    print(path)
