"""ctypes binding of libmsda_b200.so (include/msda_b200.h).  No fallback: if the library is missing or a
call fails, this raises."""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# MSDA_B200_LIB: load another build of the library instead (A/B timing of kernel revisions); never built automatically
LIB_PATH = os.environ.get("MSDA_B200_LIB") or os.path.join(_HERE, "libmsda_b200.so")
_lock = threading.Lock()
_lib = None
ABI_VERSION = 4      # include/msda_b200.h MSDA_B200_ABI_VERSION; bumped whenever the exported symbol list changes

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
    "msda_b200_shape_mismatch_epoch",
    "msda_b200_pair_value_bf16",
    "msda_b200_forward_paired_bf16",
    "msda_b200_forward_fused_paired_bf16",
    "msda_b200_small_mha_f32",
    "msda_b200_point_pos_embed_f32",
    "msda_b200_refine_points_f32",
    "msda_b200_resample_u8_hwc",
    "msda_b200_jpeg_info",
    "msda_b200_jpeg_decode_u8",
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
    """Load the CUDA library.  Builds it first if it is absent, or stale and nvcc is available (``build.build()`` is a
    no-op on a fresh library and serialises concurrent builders -- one rank per GPU -- with a file lock)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        if os.environ.get("MSDA_B200_LIB"):
            pass
        elif not os.path.exists(LIB_PATH) or os.environ.get("MSDA_B200_REBUILD") == "1":
            _build.build(force=os.environ.get("MSDA_B200_REBUILD") == "1")
        elif _build.have_nvcc():
            _build.build()                              # rebuilds only if a source is newer than the library
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
        L.msda_b200_pair_value_bf16.restype = ci
        L.msda_b200_pair_value_bf16.argtypes = [vp, ci, vp, vp, ci, ci, ci, ci, ci, vp, vp]
        L.msda_b200_forward_paired_bf16.restype = ci
        L.msda_b200_forward_paired_bf16.argtypes = [vp] * 5 + [ci] * 7 + [vp, vp]
        L.msda_b200_forward_fused_paired_bf16.restype = ci
        L.msda_b200_forward_fused_paired_bf16.argtypes = [vp, vp, vp, vp, ci, vp, vp] + [ci] * 7 + [vp, vp]
        L.msda_b200_small_mha_f32.restype = ci
        L.msda_b200_small_mha_f32.argtypes = [vp, vp, vp, ci, vp, ci, ci, ci, ci, ci, ctypes.c_longlong, ctypes.c_longlong, vp]
        L.msda_b200_point_pos_embed_f32.restype = ci
        L.msda_b200_point_pos_embed_f32.argtypes = [vp, vp, ci, vp, ctypes.c_longlong, ci, ci, vp, vp]
        L.msda_b200_refine_points_f32.restype = ci
        L.msda_b200_refine_points_f32.argtypes = [vp, vp, ctypes.c_longlong, ctypes.c_float, vp, vp]
        L.msda_b200_jpeg_info.restype = ci
        L.msda_b200_jpeg_info.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci)]
        L.msda_b200_jpeg_decode_u8.restype = ci
        L.msda_b200_jpeg_decode_u8.argtypes = [vp, ctypes.c_size_t, ci, vp, ci, ci, vp]
        L.msda_b200_resample_u8_hwc.restype = ci
        L.msda_b200_resample_u8_hwc.argtypes = [vp, ci, ci, ci, ci, ci, vp, vp, ci, ci, vp, vp]
        L.msda_b200_shape_mismatch_epoch.restype = ci
        L.msda_b200_shape_mismatch_epoch.argtypes = []
        if L.msda_b200_abi_version() != ABI_VERSION:
            raise MSDAError("libmsda_b200.so ABI version %d, expected %d (stale library: rebuild with "
                            "python -m gomatching_b200.build --force)" % (L.msda_b200_abi_version(), ABI_VERSION))
        _lib = L
    return _lib


# Measurement hooks (bench.py): ``calls`` counts the kernel-launching C-ABI calls made through this module (every one
# of them goes through ``check``; each enqueues at least one of this library's kernels); ``event_log``, when set to a
# list, makes the sampler entry points bracket their launch with CUDA events on the launching stream and append
# ``(tag, start_event, stop_event)`` -- how bench.py times the sampler live inside the model's forward.
calls = 0
event_log = None


def check(rc: int, what: str = "msda_b200") -> None:
    global calls
    calls += 1
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


# ---- host-side level geometry for the TMA window kernel, keyed by VALUE ------------------------------------------------
# The operator API hands over spatial_shapes / level_start_index as device tensors, and the reference builds fresh ones
# for every frame (deformable_transformer.py:169), so an identity-keyed cache would pay a device->host read per frame.
# Here the key is what identifies a pyramid cheaply on the host -- (device, S, L) -- and the entry is learned with ONE
# device->host read the first time the key is seen (the reference's own `assert` at ms_deform_attn.py:131 pays such a
# sync on every call).  Two pyramids can share S (portrait vs landscape), so every launch validates the entry against
# the device tensors in-kernel (csrc/msda_forward_pipelined.cu shape guard): a wrong entry costs speed, never
# correctness, and the mismatch epoch the kernels leave in pinned host memory drops the cache on a later call.
_window_geometry = {}
_window_seen_epoch = 0
window_stats = {"learned": 0, "dropped": 0, "from_list": 0}


def window_geometry(spatial_shapes, level_start_index, S: int, L: int, shapes_list=None) -> bool:
    """Set the calling thread's window-kernel geometry hint for this pyramid.  False if it is not known and cannot be
    learned right now (CUDA graph capture in progress): the caller then keeps the register-gather kernel."""
    global _window_seen_epoch
    import torch
    L_ = lib()
    ep = int(L_.msda_b200_shape_mismatch_epoch())
    if ep != _window_seen_epoch:
        _window_seen_epoch = ep
        window_stats["dropped"] += len(_window_geometry)
        _window_geometry.clear()
    dev = spatial_shapes.device.index if spatial_shapes.is_cuda else -1
    key = (dev, int(S), int(L))
    hit = _window_geometry.get(key)
    if shapes_list is not None:
        flat = [int(v) for hw in shapes_list for v in hw]
        if hit is None or hit[2] != flat:
            sh = torch.tensor(flat, dtype=torch.int64).view(-1, 2)
            if sh.shape[0] != L or int(sh.prod(1).sum()) != S:
                raise ValueError("spatial_shapes_list %r does not describe %d levels with %d pixels" % (shapes_list, L, S))
            ls = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1])).contiguous()
            hit = _window_geometry[key] = (sh.contiguous(), ls, flat)
            window_stats["from_list"] += 1
    elif hit is None:
        if torch.cuda.is_current_stream_capturing():
            return False
        sh = spatial_shapes.detach().to("cpu", copy=True).to(dtype=torch.int64).contiguous()
        ls = level_start_index.detach().to("cpu", copy=True).to(dtype=torch.int64).contiguous()
        if len(_window_geometry) > 256:
            _window_geometry.clear()
        hit = _window_geometry[key] = (sh, ls, [int(v) for v in sh.flatten()])
        window_stats["learned"] += 1
    L_.msda_b200_staged_set_host_shapes(hit[0].data_ptr(), hit[1].data_ptr(), int(hit[1].numel()))
    return True


def auto_window_tuning(value, spatial_shapes, level_start_index, Lq: int, tuning, shapes_list=None):
    """The launch plan for one sampler call.  An explicit ``tuning`` wins (modes 4 / 5 get their geometry hint here);
    without one, fp32 encoder self-attention at the DeepSolo shape (Lq == S, D = 32, L = 4, P = 4) runs the TMA
    window kernel (mode 5) whenever the pyramid's geometry is known on the host."""
    import torch
    N, S, M, D = value.shape
    L = int(spatial_shapes.shape[0])
    mode = tuning_mode(tuning)
    if tuning is not None:
        if mode in (4, 5):
            if shapes_list is not None or os.environ.get("MSDA_B200_IDENTITY_HINT") != "1":
                window_geometry(spatial_shapes, level_start_index, S, L, shapes_list)
            else:
                staged_shape_hint(spatial_shapes, level_start_index)
        return tuning
    if (value.dtype == torch.float32 and Lq == S and D == 32 and L == 4 and os.environ.get("MSDA_B200_NO_WINDOW") != "1"
            and window_geometry(spatial_shapes, level_start_index, S, L, shapes_list)):
        # the window kernel gives each CTA (one per SM) a contiguous run of (frame, head, 8x16 tile) items and needs a
        # couple of dozen of them to amortise its pipeline ramp: 720p, fused: 89.9 vs 77.3 us at 1 frame per launch,
        # 150 vs 143 at 2, 522 vs 544 at 8 (profiles/r02_sweep_encoder_f1_f2.log)
        h0, w0 = _window_geometry[(spatial_shapes.device.index, int(S), L)][2][:2]
        items = N * M * ((h0 + 7) // 8) * ((w0 + 15) // 16)
        if items >= WINDOW_MIN_ITEMS_PER_SM * _sm_count(value.device):
            return _MODE5
    return None


WINDOW_MIN_ITEMS_PER_SM = 24
_sm_counts = {}


def _sm_count(device) -> int:
    n = _sm_counts.get(device.index)
    if n is None:
        n = _sm_counts[device.index] = int(lib().msda_b200_sm_count())
    return n


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


_MODE5 = {"mode": 5}
