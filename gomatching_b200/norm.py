"""Residual add + LayerNorm as one HBM pass (csrc/add_layernorm.cu).

Reference: ``src = src + dropout(src2); src = norm(src)`` after every attention / feed-forward block
(third_party/adet/layers/deformable_transformer.py:251-252, :272-273) -- an eager add plus torch's LayerNorm kernel.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import _native

__all__ = ["add_layernorm", "add_layernorm_supported"]


def add_layernorm_supported(x: torch.Tensor, norm: nn.LayerNorm) -> bool:
    c = x.shape[-1]
    return (x.is_cuda and x.dtype == torch.float32 and len(norm.normalized_shape) == 1 and norm.normalized_shape[0] == c
            and c % 128 == 0 and c <= 1024 and (norm.weight is None or norm.weight.dtype == torch.float32))


def add_layernorm(x: torch.Tensor, y: "torch.Tensor | None", norm: nn.LayerNorm) -> torch.Tensor:
    """``norm(x + y)`` (``norm(x)`` if y is None) for fp32 CUDA tensors normalised over the last dimension."""
    if not x.is_cuda:
        raise RuntimeError("add_layernorm: Not implemented on the CPU")
    if not add_layernorm_supported(x, norm) or (y is not None and (y.shape != x.shape or y.dtype != x.dtype)):
        return norm(x if y is None else x + y)
    c = x.shape[-1]
    xc = x.contiguous()
    yc = y.contiguous() if y is not None else None
    out = torch.empty_like(xc)
    w = norm.weight.detach().contiguous() if norm.weight is not None else None
    b = norm.bias.detach().contiguous() if norm.bias is not None else None
    with torch.cuda.device(x.device):
        rc = _native.lib().msda_b200_add_layernorm_f32(
            xc.data_ptr(), yc.data_ptr() if yc is not None else None, w.data_ptr() if w is not None else None,
            b.data_ptr() if b is not None else None, float(norm.eps), xc.numel() // c, c, out.data_ptr(),
            torch.cuda.current_stream().cuda_stream)
    _native.check(rc, "msda_b200_add_layernorm_f32")
    return out
