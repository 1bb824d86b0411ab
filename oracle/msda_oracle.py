"""TEST INFRASTRUCTURE -- Python face of the CPU oracle.  NOT part of the product.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  ``gomatching_b200/`` never does.

Two independent restatements of the reference's MSDeformAttn core live here:

* ``forward_f32 / forward_f64 / forward_bf16 / sample_index`` -- ctypes wrappers over
  ``oracle/msda_oracle.c``: the reference CUDA kernel's arithmetic, operation by operation
  (``ms_deform_im2col_cuda.cuh:33-84, 237-299``).  This is the *index oracle* and the bit-level
  output oracle.
* ``core_gridsample`` -- the reference's own CPU path restated: per-level ``F.grid_sample``
  (bilinear, zero padding, ``align_corners=False``) on ``2*loc-1`` followed by the attention-weighted
  sum (``third_party/adet/layers/ms_deform_attn.py:40-60``).  This is the *reference CPU
  implementation* that ``bench.py`` times as ``cpu_baseline`` (kind "port": the reference's Python
  file itself cannot travel to the GPU box).

``module_forward`` restates ``MSDeformAttn.forward`` (``ms_deform_attn.py:117-156``) on top of either.

Parity pin: both are checked in ``tests/test_oracle.py`` against fixtures produced by importing the
reference's real ``ms_deform_attn_core_pytorch`` / ``MSDeformAttn`` / ``DeformableTransformer`` in the
build container (``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmsda_oracle.so")
_lib = None


class IndexRecord(ctypes.Structure):
    _fields_ = [
        ("h_low", ctypes.c_int32),
        ("w_low", ctypes.c_int32),
        ("in_range", ctypes.c_int32),
        ("corner_mask", ctypes.c_int32),
        ("level_offset", ctypes.c_int64),
    ]


INDEX_DTYPE = np.dtype(
    [("h_low", "<i4"), ("w_low", "<i4"), ("in_range", "<i4"), ("corner_mask", "<i4"), ("level_offset", "<i8")]
)


def build(force: bool = False) -> str:
    """Compile oracle/msda_oracle.c with gcc (oracle/Makefile).  Building the checker is not using it."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "msda_oracle.c")
    ):
        subprocess.run(["make", "-C", _HERE, "oracle"], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.msda_oracle_version.restype = ctypes.c_int
        _lib.msda_oracle_max_threads.restype = ctypes.c_int
    return _lib


def max_threads() -> int:
    return int(lib().msda_oracle_max_threads())


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a), dtype=dtype)


def _dims(value, loc):
    N, S, M, D = value.shape
    N2, Lq, M2, L, P, two = loc.shape
    assert (N2, M2, two) == (N, M, 2), "value/loc shape mismatch"
    return N, S, M, D, L, Lq, P


def level_start_index(shapes: Sequence[Sequence[int]]) -> np.ndarray:
    """deformable_transformer.py:170 -- cat(0, cumsum(H*W)[:-1])"""
    hw = np.asarray(shapes, dtype=np.int64)
    areas = hw[:, 0] * hw[:, 1]
    return np.concatenate([[0], np.cumsum(areas)[:-1]]).astype(np.int64)


def forward_f32(value, shapes, lsi, loc, attn) -> np.ndarray:
    value, loc, attn = _c(value, np.float32), _c(loc, np.float32), _c(attn, np.float32)
    shapes, lsi = _c(shapes, np.int64), _c(lsi, np.int64)
    N, S, M, D, L, Lq, P = _dims(value, loc)
    out = np.empty((N, Lq, M * D), dtype=np.float32)
    rc = lib().msda_oracle_forward_f32(_p(value), _p(shapes), _p(lsi), _p(loc), _p(attn), N, S, M, D, L, Lq, P, _p(out))
    assert rc == 0
    return out


def forward_f64(value, shapes, lsi, loc, attn) -> np.ndarray:
    value, loc, attn = _c(value, np.float64), _c(loc, np.float64), _c(attn, np.float64)
    shapes, lsi = _c(shapes, np.int64), _c(lsi, np.int64)
    N, S, M, D, L, Lq, P = _dims(value, loc)
    out = np.empty((N, Lq, M * D), dtype=np.float64)
    rc = lib().msda_oracle_forward_f64(_p(value), _p(shapes), _p(lsi), _p(loc), _p(attn), N, S, M, D, L, Lq, P, _p(out))
    assert rc == 0
    return out


def forward_exact(value, shapes, lsi, loc, attn) -> np.ndarray:
    """fp32 inputs evaluated in double without any fusion: the tolerance yardstick."""
    value, loc, attn = _c(value, np.float32), _c(loc, np.float32), _c(attn, np.float32)
    shapes, lsi = _c(shapes, np.int64), _c(lsi, np.int64)
    N, S, M, D, L, Lq, P = _dims(value, loc)
    out = np.empty((N, Lq, M * D), dtype=np.float64)
    rc = lib().msda_oracle_forward_exact(_p(value), _p(shapes), _p(lsi), _p(loc), _p(attn), N, S, M, D, L, Lq, P, _p(out))
    assert rc == 0
    return out


def f32_to_bf16_bits(x) -> np.ndarray:
    """round-to-nearest-even fp32 -> bf16 bit pattern (uint16), same rule as __float2bfloat16_rn."""
    u = _c(x, np.float32).view(np.uint32).astype(np.uint64)
    lsb = (u >> 16) & 1
    r = ((u + 0x7FFF + lsb) >> 16).astype(np.uint16)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    r[nan] = 0x7FFF
    return r


def bf16_bits_to_f32(b) -> np.ndarray:
    return (_c(b, np.uint16).astype(np.uint32) << 16).view(np.float32)


def forward_bf16(value_bits, shapes, lsi, loc, attn) -> np.ndarray:
    """value/out as bf16 bit patterns (uint16); loc/attn fp32; fp32 arithmetic, one final rounding."""
    value_bits, loc, attn = _c(value_bits, np.uint16), _c(loc, np.float32), _c(attn, np.float32)
    shapes, lsi = _c(shapes, np.int64), _c(lsi, np.int64)
    N, S, M, D, L, Lq, P = _dims(value_bits, loc)
    out = np.empty((N, Lq, M * D), dtype=np.uint16)
    rc = lib().msda_oracle_forward_bf16(_p(value_bits), _p(shapes), _p(lsi), _p(loc), _p(attn), N, S, M, D, L, Lq, P, _p(out))
    assert rc == 0
    return out


def sample_index(loc, shapes, lsi, M: int, D: int) -> np.ndarray:
    """structured array (N,Lq,M,L,P) of (h_low, w_low, in_range, corner_mask, level_offset)."""
    loc = _c(loc, np.float32)
    shapes, lsi = _c(shapes, np.int64), _c(lsi, np.int64)
    N, Lq, M2, L, P, _ = loc.shape
    assert M2 == M
    out = np.zeros((N, Lq, M, L, P), dtype=INDEX_DTYPE)
    rc = lib().msda_oracle_sample_index_f32(_p(loc), _p(shapes), _p(lsi), N, Lq, M, D, L, P, _p(out))
    assert rc == 0
    return out


def locations(ref, off, shapes, n_points: int | None = None) -> np.ndarray:
    """ms_deform_attn.py:141-147.  ref (N,Lq,L,2|4), off (N,Lq,M,L,P,2) -> loc (N,Lq,M,L,P,2)."""
    ref, off = _c(ref, np.float32), _c(off, np.float32)
    shapes = _c(shapes, np.int64)
    N, Lq, M, L, P, _ = off.shape
    assert ref.shape[:3] == (N, Lq, L)
    loc = np.empty_like(off)
    rc = lib().msda_oracle_locations_f32(_p(ref), int(ref.shape[-1]), _p(off), _p(shapes), N, Lq, M, L, P, _p(loc))
    if rc != 0:
        raise ValueError(
            "Last dim of reference_points must be 2 or 4, but get {} instead.".format(ref.shape[-1])
        )
    return loc


def softmax(logits) -> np.ndarray:
    logits = _c(logits, np.float32)
    out = np.empty_like(logits)
    cols = logits.shape[-1]
    rc = lib().msda_oracle_softmax_f32(_p(logits), logits.size // cols, cols, _p(out))
    assert rc == 0
    return out


# ------------------------------------------------------------------------------------------------
# The reference's CPU path, restated (ms_deform_attn.py:40-60).  torch is used exactly as the
# reference uses it: F.grid_sample does the bilinear gather.
# ------------------------------------------------------------------------------------------------
def core_gridsample(value, shapes: Sequence[Sequence[int]], loc, attn):
    import torch
    import torch.nn.functional as F

    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    hw = [(int(h), int(w)) for h, w in shapes]
    per_level = value.split([h * w for h, w in hw], dim=1)
    grids = 2 * loc - 1                                           # [0,1] -> [-1,1]           (:46)
    sampled = []
    for lvl, (h, w) in enumerate(hw):
        # (N, h*w, M, D) -> (N*M, D, h, w)                                                    (:50)
        v = per_level[lvl].flatten(2).transpose(1, 2).reshape(N * M, D, h, w)
        # (N, Lq, M, P, 2) -> (N*M, Lq, P, 2)                                                 (:52)
        g = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    a = attn.transpose(1, 2).reshape(N * M, 1, Lq, L * P)         # (:57)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * a).sum(-1).view(N, M * D, Lq)
    return out.transpose(1, 2).contiguous()


def module_forward(params: dict, query, reference_points, input_flatten, shapes, lsi, padding_mask=None,
                   n_heads: int = 8, n_levels: int = 4, n_points: int = 4, core: str = "kernel"):
    """MSDeformAttn.forward (ms_deform_attn.py:117-156) on CPU tensors.

    params: state-dict style {'sampling_offsets.weight', 'sampling_offsets.bias', 'attention_weights.*',
    'value_proj.*', 'output_proj.*'}.  core = "kernel" (C restatement of the CUDA kernel) or
    "gridsample" (the reference CPU path).  Returns (output, sampling_locations, attention_weights).
    """
    import torch
    import torch.nn.functional as F

    N, Lq, C = query.shape
    _, S, _ = input_flatten.shape
    shapes_t = torch.as_tensor(np.asarray(shapes), dtype=torch.int64)
    assert int((shapes_t[:, 0] * shapes_t[:, 1]).sum()) == S                                   # (:131)
    value = F.linear(input_flatten, params["value_proj.weight"], params["value_proj.bias"])
    if padding_mask is not None:
        value = value.masked_fill(padding_mask[..., None], float(0))
    value = value.view(N, S, n_heads, C // n_heads)
    off = F.linear(query, params["sampling_offsets.weight"], params["sampling_offsets.bias"]).view(
        N, Lq, n_heads, n_levels, n_points, 2)
    aw = F.linear(query, params["attention_weights.weight"], params["attention_weights.bias"]).view(
        N, Lq, n_heads, n_levels * n_points)
    aw = F.softmax(aw, -1).view(N, Lq, n_heads, n_levels, n_points)
    if reference_points.shape[-1] == 2:
        normalizer = torch.stack([shapes_t[..., 1], shapes_t[..., 0]], -1)
        loc = reference_points[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
    elif reference_points.shape[-1] == 4:
        loc = reference_points[:, :, None, :, None, :2] + off / n_points * reference_points[:, :, None, :, None, 2:] * 0.5
    else:
        raise ValueError(
            "Last dim of reference_points must be 2 or 4, but get {} instead.".format(reference_points.shape[-1]))
    if core == "kernel":
        o = torch.from_numpy(forward_f32(value.numpy(), shapes_t.numpy(), np.asarray(lsi), loc.numpy(), aw.numpy()))
    else:
        o = core_gridsample(value, shapes_t.tolist(), loc, aw)
    out = F.linear(o, params["output_proj.weight"], params["output_proj.bias"])
    return out, loc, aw
