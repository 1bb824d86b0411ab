"""ctypes binding of libmsda_b200.so (include/msda_b200.h).  No fallback: if the library is missing or a
call fails, this raises."""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmsda_b200.so")
_lock = threading.Lock()
_lib = None

# every symbol include/msda_b200.h declares; tests/test_abi.py checks header <-> list <-> .so agree
SYMBOLS = (
    "msda_b200_abi_version",
    "msda_b200_error_string",
    "msda_b200_sm_count",
    "msda_b200_variant_count",
    "msda_b200_forward_f32",
    "msda_b200_forward_bf16",
    "msda_b200_forward_f64",
    "msda_b200_forward_f32_ex",
    "msda_b200_forward_bf16_ex",
    "msda_b200_forward_fused_f32",
    "msda_b200_forward_fused_bf16",
    "msda_b200_linear_split_weight_f32",
    "msda_b200_linear_f32",
    "msda_b200_linear_relu_f32",
    "msda_b200_linear_set_trace",
    "msda_b200_forward_fused_pitched_f32",
    "msda_b200_forward_fused_pitched_bf16",
    "msda_b200_locations_softmax_f32",
    "msda_b200_sample_index_f32",
    "msda_b200_backward_f32",
    "msda_b200_host_ctx_create",
    "msda_b200_host_ctx_destroy",
    "msda_b200_forward_f32_host",
    "msda_b200_frames_u8_to_chw_f32",
    "msda_b200_staged_set_host_shapes",
    "msda_b200_add_layernorm_f32",
)


class Tuning(ctypes.Structure):
    """msda_b200_tuning_t"""
    _fields_ = [
        ("mode", ctypes.c_int32),
        ("tile_h", ctypes.c_int32),
        ("tile_w", ctypes.c_int32),
        ("tile_q", ctypes.c_int32),
        ("ctas_per_sm", ctypes.c_int32),
        ("variant", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 2),
    ]


MODE_AUTO, MODE_LINEAR, MODE_PYRAMID, MODE_GENERIC = 0, 1, 2, 3


class MSDAError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load the CUDA library.  Builds it first if nvcc is available and the .so is stale or absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH) or os.environ.get("MSDA_B200_REBUILD") == "1":
            from . import build as _build
            _build.build()
        if not os.path.exists(LIB_PATH):
            raise MSDAError("libmsda_b200.so is missing and could not be built; there is no CPU/PyTorch fallback")
        L = ctypes.CDLL(LIB_PATH)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        L.msda_b200_abi_version.restype = ci
        L.msda_b200_error_string.restype = ctypes.c_char_p
        L.msda_b200_error_string.argtypes = [ci]
        L.msda_b200_sm_count.restype = ci
        L.msda_b200_variant_count.restype = ci
        core = [vp, vp, vp, vp, vp] + [ci] * 7 + [vp, vp]
        for name in ("msda_b200_forward_f32", "msda_b200_forward_bf16", "msda_b200_forward_f64"):
            getattr(L, name).restype = ci
            getattr(L, name).argtypes = core
        for name in ("msda_b200_forward_f32_ex", "msda_b200_forward_bf16_ex"):
            getattr(L, name).restype = ci
            getattr(L, name).argtypes = core + [ctypes.POINTER(Tuning)]
        fused = [vp, vp, vp, vp, ci, vp, vp] + [ci] * 7 + [vp, vp, ctypes.POINTER(Tuning)]
        for name in ("msda_b200_forward_fused_f32", "msda_b200_forward_fused_bf16"):
            getattr(L, name).restype = ci
            getattr(L, name).argtypes = fused
        pitched = [vp, vp, vp, vp, ci, vp, ci, vp, ci] + [ci] * 7 + [vp, vp, ctypes.POINTER(Tuning)]
        for name in ("msda_b200_forward_fused_pitched_f32", "msda_b200_forward_fused_pitched_bf16"):
            getattr(L, name).restype = ci
            getattr(L, name).argtypes = pitched
        L.msda_b200_locations_softmax_f32.restype = ci
        L.msda_b200_locations_softmax_f32.argtypes = [vp, vp, ci, vp, vp] + [ci] * 6 + [vp, vp, vp]
        L.msda_b200_sample_index_f32.restype = ci
        L.msda_b200_sample_index_f32.argtypes = [vp, vp, vp] + [ci] * 6 + [vp, vp]
        L.msda_b200_backward_f32.restype = ci
        L.msda_b200_backward_f32.argtypes = [vp] * 6 + [ci] * 7 + [vp, vp, vp, vp]
        L.msda_b200_host_ctx_create.restype = ci
        L.msda_b200_host_ctx_create.argtypes = [ctypes.POINTER(vp), ci]
        L.msda_b200_host_ctx_destroy.restype = None
        L.msda_b200_host_ctx_destroy.argtypes = [vp]
        L.msda_b200_forward_f32_host.restype = ci
        L.msda_b200_forward_f32_host.argtypes = [vp] * 6 + [ci] * 7 + [vp]
        L.msda_b200_linear_split_weight_f32.restype = ci
        L.msda_b200_linear_split_weight_f32.argtypes = [vp, ci, ci, vp, vp, vp]
        L.msda_b200_linear_f32.restype = ci
        L.msda_b200_linear_f32.argtypes = [vp, ci, vp, vp, vp, vp, ci, ci, ci, vp, ci, vp]
        L.msda_b200_linear_relu_f32.restype = ci
        L.msda_b200_linear_relu_f32.argtypes = [vp, ci, vp, vp, vp, ci, ci, ci, vp, ci, vp]
        L.msda_b200_linear_set_trace.restype = None
        L.msda_b200_linear_set_trace.argtypes = [vp]
        L.msda_b200_add_layernorm_f32.restype = ci
        L.msda_b200_add_layernorm_f32.argtypes = [vp, vp, vp, vp, ctypes.c_float, ctypes.c_longlong, ci, vp, vp]
        L.msda_b200_staged_set_host_shapes.restype = None
        L.msda_b200_staged_set_host_shapes.argtypes = [vp, vp, ci]
        L.msda_b200_frames_u8_to_chw_f32.restype = ci
        L.msda_b200_frames_u8_to_chw_f32.argtypes = [vp, ci, ci, ci, ci, vp, vp, ci, ci, vp, vp]
        if L.msda_b200_abi_version() != 1:
            raise MSDAError("libmsda_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc: int, what: str = "msda_b200") -> None:
    if rc != 0:
        msg = lib().msda_b200_error_string(int(rc))
        raise MSDAError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


_shape_hint_cache = {}


def staged_shape_hint(spatial_shapes, level_start_index) -> None:
    """tuning modes 4 / 5: hand the level geometry to the library as host arrays so its window fills can use TMA.
    One device->host read per distinct shapes tensor, none afterwards.  The cache entry keeps the two tensors alive,
    so their identity (and storage pointer) cannot be recycled for another pyramid while the entry exists, and an
    in-place modification is caught by the version counters.  The library keeps the hint per calling thread, and it is
    set immediately before every mode-4 / mode-5 launch."""
    import torch
    key = (id(spatial_shapes), id(level_start_index))
    hit = _shape_hint_cache.get(key)
    if (hit is None or hit[0] is not spatial_shapes or hit[1] is not level_start_index
            or hit[2] != (spatial_shapes._version, level_start_index._version)):
        if len(_shape_hint_cache) > 64:
            _shape_hint_cache.clear()
        sh = spatial_shapes.detach().to("cpu", copy=True).to(dtype=torch.int64).contiguous()
        ls = level_start_index.detach().to("cpu", copy=True).to(dtype=torch.int64).contiguous()
        hit = _shape_hint_cache[key] = (spatial_shapes, level_start_index,
                                        (spatial_shapes._version, level_start_index._version), sh, ls)
    sh, ls = hit[3], hit[4]
    lib().msda_b200_staged_set_host_shapes(sh.data_ptr(), ls.data_ptr(), int(ls.numel()))


def tuning_mode(tuning) -> int:
    if tuning is None:
        return 0
    return int(tuning.mode) if isinstance(tuning, Tuning) else int(dict(tuning).get("mode", 0))


def make_tuning(tuning) -> "ctypes.POINTER(Tuning) | None":
    """dict / Tuning / None -> pointer usable in the *_ex calls"""
    if tuning is None:
        return None
    if isinstance(tuning, Tuning):
        return ctypes.pointer(tuning)
    t = Tuning()
    for k, v in dict(tuning).items():
        if k == "force_v1":            # reserved[0] = 1: run the runtime-L*P tiled kernel instead of the specialised one
            t.reserved[0] = int(v)
        elif k == "walk":              # reserved[1] = 1: contiguous raster tile walk per CTA (diagnostic) instead of the strided one
            t.reserved[1] = int(v)
        else:
            setattr(t, k, int(v))
    return ctypes.pointer(t)
