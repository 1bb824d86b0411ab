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
SOURCES = ["msda_forward.cu", "msda_forward_fast.cu", "msda_forward_staged.cu", "msda_forward_pipelined.cu", "msda_backward.cu", "msda_api.cu", "proj_gemm.cu", "frame_batcher.cu", "add_layernorm.cu"]
HEADERS = ["msda_device.cuh", "msda_fast_common.cuh", "tma_common.cuh", "msda_launch.h", os.path.join("..", "..", "include", "msda_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    # no --use_fast_math: division, expf and the explicit __f*_rn chain must stay IEEE (index contract)
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libmsda_b200.so cannot be built (there is no CPU fallback)")
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(CSRC, s.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", o]
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
    tmp = LIB + ".tmp"
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    os.replace(tmp, LIB)
    for o in objs:
        try:
            os.remove(o)
        except OSError:
            pass
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
