"""TEST INFRASTRUCTURE -- ctypes front end of oracle/jpeg_oracle.cpp (CPU restatement of libjpeg's baseline decode).
Only tests/ may import this."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libjpeg_oracle.so")
        src = os.path.join(_HERE, "jpeg_oracle.cpp")
        hdr = os.path.join(_HERE, "..", "gomatching_b200", "csrc", "jpeg_entropy.h")
        if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            subprocess.run(["make", "-C", _HERE, "jpeg"], check=True, stdout=subprocess.DEVNULL)
        _lib = ctypes.CDLL(path)
        _lib.jpeg_oracle_decode.restype = ctypes.c_int
        _lib.jpeg_oracle_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int),
                                            ctypes.POINTER(ctypes.c_int), ctypes.c_void_p, ctypes.c_size_t]
    return _lib


def decode_rgb(data: bytes) -> np.ndarray:
    """(H, W, 3) uint8 RGB, or raises ValueError(code): 1 truncated, 2 corrupt, 3 unsupported."""
    w, h = ctypes.c_int(0), ctypes.c_int(0)
    rc = lib().jpeg_oracle_decode(data, len(data), ctypes.byref(w), ctypes.byref(h), None, 0)
    if rc != 0:
        raise ValueError(rc)
    out = np.empty((h.value, w.value, 3), dtype=np.uint8)
    rc = lib().jpeg_oracle_decode(data, len(data), ctypes.byref(w), ctypes.byref(h), out.ctypes.data, out.nbytes)
    if rc != 0:
        raise ValueError(rc)
    return out
