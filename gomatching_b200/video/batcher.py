"""Frame batcher: uint8 HWC frames -> the normalised, padded float32 batch the spotter's backbone takes.

Reference, per frame (N = 1 per forward):
    GoMBatchPredictor.__call__   gomatching/text_track_visualizer.py:313-321   optional RGB flip ``x[:, :, ::-1]``,
                                  ``torch.as_tensor(x.astype("float32").transpose(2, 0, 1))`` on the host
    GoMatching.preprocess_image  gomatching/modeling/meta_arch/gom_lstmatcher.py:159-170  ``(x - mean) / std`` on the
                                  device, ``ImageList.from_tensors`` (zero padding to the size divisibility)
Here the uint8 frames go to the device as they are (a quarter of the PCIe bytes) and one CUDA kernel
(``csrc/frame_batcher.cu``) writes the batch; values are bit-identical to the eager ops.  No CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Sequence

import torch

from .. import _native

__all__ = ["padded_size", "batch_frames"]


def padded_size(height: int, width: int, size_divisibility: int = 0):
    """(Hp, Wp) of ``ImageList.from_tensors``: each side rounded up to a multiple of ``size_divisibility`` (> 1)."""
    d = int(size_divisibility)
    if d > 1:
        return (height + d - 1) // d * d, (width + d - 1) // d * d
    return int(height), int(width)


def batch_frames(frames_u8: torch.Tensor, pixel_mean: Sequence[float], pixel_std: Sequence[float],
                 flip_channels: bool = False, size_divisibility: int = 0, out: "torch.Tensor | None" = None) -> torch.Tensor:
    """frames_u8: CUDA uint8 (N, H, W, 3) or (H, W, 3).  Returns float32 (N, 3, Hp, Wp).

    ``flip_channels`` reverses the channel order first (the predictor's ``input_format == "RGB"`` branch);
    ``pixel_mean`` / ``pixel_std`` are in the order of the OUTPUT channels, like ``cfg.MODEL.PIXEL_MEAN``."""
    if frames_u8.dim() == 3:
        frames_u8 = frames_u8.unsqueeze(0)
    if frames_u8.dim() != 4 or frames_u8.shape[-1] != 3 or frames_u8.dtype != torch.uint8:
        raise ValueError("frames must be uint8 (N, H, W, 3), got %s %s" % (tuple(frames_u8.shape), frames_u8.dtype))
    if not frames_u8.is_cuda:
        raise RuntimeError("batch_frames: Not implemented on the CPU (frames must be CUDA tensors; there is no CPU path)")
    if len(pixel_mean) != 3 or len(pixel_std) != 3:
        raise ValueError("pixel_mean / pixel_std must have 3 entries")
    frames_u8 = frames_u8.contiguous()
    n, h, w, _ = frames_u8.shape
    hp, wp = padded_size(h, w, size_divisibility)
    if out is None:
        out = torch.empty((n, 3, hp, wp), dtype=torch.float32, device=frames_u8.device)
    elif tuple(out.shape) != (n, 3, hp, wp) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != frames_u8.device:
        raise ValueError("out must be a contiguous float32 (%d, 3, %d, %d) tensor on %s" % (n, hp, wp, frames_u8.device))
    mean = (ctypes.c_float * 3)(*[float(v) for v in pixel_mean])
    std = (ctypes.c_float * 3)(*[float(v) for v in pixel_std])
    with torch.cuda.device(frames_u8.device):
        rc = _native.lib().msda_b200_frames_u8_to_chw_f32(
            frames_u8.data_ptr(), n, h, w, int(bool(flip_channels)), ctypes.cast(mean, ctypes.c_void_p),
            ctypes.cast(std, ctypes.c_void_p), hp, wp, out.data_ptr(),
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _native.check(rc, "msda_b200_frames_u8_to_chw_f32")
    return out
