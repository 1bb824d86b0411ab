"""JPEG frames decoded on the device, bit-identical to the reference's frame read.

The reference's video loop reads every frame with ``read_image(path, format="BGR")`` (eval.py:324-327) -- detectron2's
``PIL.Image.open -> _apply_exif_orientation -> convert("RGB") -> numpy -> [:, :, ::-1]`` -- on the host.  ``decode_jpeg`` /
``read_image_bgr`` hand the FILE BYTES to ``msda_b200_jpeg_decode_u8`` (csrc/jpeg_decode.cu): Huffman decoding on the
calling host thread, inverse DCT / chroma upsampling / colour conversion on the device, the same integer arithmetic as
Pillow's libjpeg-turbo (tests/test_jpeg_decode.py: every pixel equal to ``PIL.Image.open``).  CUDA only; formats the
kernel path does not take (progressive, CMYK, ...) raise -- there is no host fallback.
"""
from __future__ import annotations

import ctypes
import struct
from typing import Optional, Tuple, Union

import torch

from .. import _native

__all__ = ["decode_jpeg", "jpeg_size", "read_image_bgr", "exif_orientation"]

_Bytes = Union[bytes, bytearray, memoryview]


def _buffer(data: _Bytes):
    if isinstance(data, bytes):
        return data, len(data)
    mv = memoryview(data).cast("B")
    return (ctypes.c_ubyte * len(mv)).from_buffer_copy(mv), len(mv)


def jpeg_size(data: _Bytes) -> Tuple[int, int, int]:
    """(height, width, components) from the SOF marker."""
    buf, n = _buffer(data)
    w, h, c = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    _native.check(_native.lib().msda_b200_jpeg_info(buf, n, ctypes.byref(w), ctypes.byref(h), ctypes.byref(c)),
                  "msda_b200_jpeg_info")
    return h.value, w.value, c.value


def decode_jpeg(data: _Bytes, device, bgr: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """uint8 (H, W, 3) on ``device`` -- B,G,R like ``read_image(..., format="BGR")`` (default) or R,G,B."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("Not implemented on the CPU")
    buf, n = _buffer(data)
    h, w, _ = jpeg_size(data)
    if out is None:
        out = torch.empty((h, w, 3), dtype=torch.uint8, device=device)
    elif out.shape != (h, w, 3) or out.dtype != torch.uint8 or not out.is_contiguous() or out.device != device:
        raise ValueError("decode_jpeg: out must be a contiguous uint8 (%d, %d, 3) tensor on %s" % (h, w, device))
    with torch.cuda.device(device):
        _native.check(_native.lib().msda_b200_jpeg_decode_u8(buf, n, 1 if bgr else 0, out.data_ptr(), w, h,
                                                             torch.cuda.current_stream(device).cuda_stream),
                      "msda_b200_jpeg_decode_u8")
    return out


def exif_orientation(data: _Bytes) -> int:
    """EXIF orientation tag (0x0112) of a JPEG, 1 if absent -- what detectron2's _apply_exif_orientation reads."""
    mv = bytes(memoryview(data)[:65536 * 2])
    pos = 2
    while pos + 4 <= len(mv) and mv[pos] == 0xFF:
        marker, seglen = mv[pos + 1], struct.unpack(">H", mv[pos + 2:pos + 4])[0]
        if marker == 0xDA or marker == 0xD9:
            break
        seg = mv[pos + 4:pos + 2 + seglen]
        if marker == 0xE1 and seg[:6] == b"Exif\x00\x00" and len(seg) >= 14:
            tiff = seg[6:]
            e = "<" if tiff[:2] == b"II" else ">"
            off = struct.unpack(e + "I", tiff[4:8])[0]
            if off + 2 <= len(tiff):
                for i in range(struct.unpack(e + "H", tiff[off:off + 2])[0]):
                    ent = tiff[off + 2 + 12 * i:off + 14 + 12 * i]
                    if len(ent) == 12 and struct.unpack(e + "H", ent[:2])[0] == 0x0112:
                        return struct.unpack(e + "H", ent[8:10])[0]
            return 1
        pos += 2 + seglen
    return 1


def read_image_bgr(path: str, device) -> torch.Tensor:
    """``read_image(path, format="BGR")`` (eval.py:327) for a JPEG file, as a device tensor.  Rotated / mirrored EXIF
    orientations get the same transpose detectron2 applies, on the device."""
    with open(path, "rb") as f:
        data = f.read()
    img = decode_jpeg(data, device, bgr=True)
    o = exif_orientation(data)
    if o in (0, 1):
        return img
    # PIL transpose methods of detectron2.data.detection_utils._apply_exif_orientation
    if o == 2:
        img = img.flip(1)                    # FLIP_LEFT_RIGHT
    elif o == 3:
        img = img.flip(0).flip(1)            # ROTATE_180
    elif o == 4:
        img = img.flip(0)                    # FLIP_TOP_BOTTOM
    elif o == 5:
        img = img.transpose(0, 1)            # TRANSPOSE
    elif o == 6:
        img = img.transpose(0, 1).flip(1)    # ROTATE_270
    elif o == 7:
        img = img.flip(0).flip(1).transpose(0, 1)   # TRANSVERSE
    elif o == 8:
        img = img.transpose(0, 1).flip(0)    # ROTATE_90
    return img.contiguous()
