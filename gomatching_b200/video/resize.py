"""Test-time frame resize, bit-identical to the reference's PIL path, on the device.

Reference: ``self.aug.get_transform(x).apply_image(x)`` with ``aug = ResizeShortestEdge([MIN_SIZE_TEST, MIN_SIZE_TEST],
MAX_SIZE_TEST)`` in both predictors (gomatching/text_track_visualizer.py:283-284, :318-319; detectron2's
``ResizeShortestEdge.get_output_shape`` and ``ResizeTransform.apply_image`` -> ``PIL.Image.resize((w, h), BILINEAR)``
for uint8 images).  Neither detectron2 nor its transform code is in the reference tree; the two pieces restated here
are (1) the output-shape rule and (2) Pillow's coefficient tables (libImaging/Resample.c ``precompute_coeffs`` +
``normalize_coeffs_8bpc``, bilinear filter, support 1, antialiased when shrinking).  The tables go to
``msda_b200_resample_u8_hwc`` (csrc/frame_resize.cu), which runs Pillow's two fixed-point passes.  Checked against
Pillow itself in tests/test_frame_resize.py (up- and down-scaling, odd sizes).  No CPU path.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np
import torch

from .. import _native

PRECISION_BITS = 32 - 8 - 2


def shortest_edge_size(h: int, w: int, short_edge: int, max_size: int) -> Tuple[int, int]:
    """detectron2 ``ResizeShortestEdge.get_output_shape``: scale the short side to ``short_edge``, cap the long side at
    ``max_size``, round half up.  Returns (new_h, new_w)."""
    size = short_edge * 1.0
    scale = size / min(h, w)
    if h < w:
        newh, neww = size, scale * w
    else:
        newh, neww = scale * h, size
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh, neww = newh * scale, neww * scale
    return int(newh + 0.5), int(neww + 0.5)


def bilinear_coeffs(in_size: int, out_size: int):
    """Pillow's ``precompute_coeffs(inSize, 0, inSize, outSize, BILINEAR)`` followed by ``normalize_coeffs_8bpc``.
    Returns (bounds int32 (out, 2) = [first, count], coeffs int32 (out, ksize), ksize).  All arithmetic in float64 with
    C's truncating ``(int)`` casts, as in Resample.c."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale                                     # bilinear: support 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)                          # (int) truncates; the operand is >= -0.5 here
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        x = np.arange(xmax, dtype=np.float64)
        w = (x + xmin - center + 0.5) * ss
        w = np.where(np.abs(w) < 1.0, 1.0 - np.abs(w), 0.0)         # bilinear_filter
        ww = 0.0
        for v in w:                                                 # same summation order as the C loop
            ww += float(v)
        if ww != 0.0:
            w = w / ww
        kk[xx, :xmax] = w
        bounds[xx] = (xmin, xmax)
    fixed = np.where(kk < 0, -0.5 + kk * (1 << PRECISION_BITS), 0.5 + kk * (1 << PRECISION_BITS))
    return bounds, np.trunc(fixed).astype(np.int32), ksize


def resample_reference(img: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """numpy emulation of the two passes (tests only: pins ``bilinear_coeffs`` against Pillow without a GPU)."""
    def one_pass(a, axis, out_size):
        bounds, kk, _ = bilinear_coeffs(a.shape[axis], out_size)
        a = np.moveaxis(a, axis, 0).astype(np.int64)
        out = np.empty((out_size,) + a.shape[1:], np.uint8)
        for o in range(out_size):
            first, cnt = bounds[o]
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[o, :cnt].astype(np.int64), a[first:first + cnt], axes=(0, 0))
            out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
        return np.moveaxis(out, 0, axis)
    out = img
    if new_w != img.shape[1]:
        out = one_pass(out, 1, new_w)                               # Pillow: horizontal pass first
    if new_h != img.shape[0]:
        out = one_pass(out, 0, new_h)
    return out


_tables = {}


def _device_tables(in_size: int, out_size: int, device):
    key = (in_size, out_size, str(device))
    hit = _tables.get(key)
    if hit is None:
        if len(_tables) > 64:
            _tables.clear()
        bounds, kk, ksize = bilinear_coeffs(in_size, out_size)
        hit = _tables[key] = (torch.from_numpy(bounds).to(device), torch.from_numpy(np.ascontiguousarray(kk)).to(device), ksize)
    return hit


def resize_frames_u8(frames: torch.Tensor, new_h: int, new_w: int) -> torch.Tensor:
    """frames: CUDA uint8 (N, H, W, C) or (H, W, C), C in {1, 3, 4}.  Returns uint8 (N, new_h, new_w, C) (same rank as the
    input), bit-identical to ``PIL.Image.fromarray(x).resize((new_w, new_h), PIL.Image.BILINEAR)`` per frame."""
    squeeze = frames.dim() == 3
    if squeeze:
        frames = frames.unsqueeze(0)
    if frames.dim() != 4 or frames.dtype != torch.uint8:
        raise ValueError("frames must be uint8 (N, H, W, C), got %s %s" % (tuple(frames.shape), frames.dtype))
    if not frames.is_cuda:
        raise RuntimeError("resize_frames_u8: Not implemented on the CPU (there is no CPU path)")
    frames = frames.contiguous()
    n, h, w, c = frames.shape
    lib = _native.lib()
    cur = frames
    with torch.cuda.device(frames.device):
        stream = torch.cuda.current_stream().cuda_stream
        if new_w != w:
            bounds, kk, ksize = _device_tables(w, new_w, frames.device)
            out = torch.empty((n, h, new_w, c), dtype=torch.uint8, device=frames.device)
            _native.check(lib.msda_b200_resample_u8_hwc(cur.data_ptr(), n, h, w, c, 1, bounds.data_ptr(), kk.data_ptr(), ksize,
                                                        new_w, out.data_ptr(), stream), "msda_b200_resample_u8_hwc")
            cur = out
        if new_h != h:
            bounds, kk, ksize = _device_tables(h, new_h, frames.device)
            out = torch.empty((n, new_h, cur.shape[2], c), dtype=torch.uint8, device=frames.device)
            _native.check(lib.msda_b200_resample_u8_hwc(cur.data_ptr(), n, h, cur.shape[2], c, 0, bounds.data_ptr(), kk.data_ptr(),
                                                        ksize, new_h, out.data_ptr(), stream), "msda_b200_resample_u8_hwc")
            cur = out
    return cur[0] if squeeze else cur
