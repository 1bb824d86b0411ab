"""Tensor-core projections bordering the sampler: nn.Linear replaced by the 3xTF32 tcgen05 GEMM of csrc/proj_gemm.cu.

Reference being replaced: the four ``nn.Linear`` calls of ``MSDeformAttn.forward``
(third_party/adet/layers/ms_deform_attn.py:133 value_proj (+ :135 masked_fill), :137 sampling_offsets,
:138 attention_weights, :153 output_proj), fp32 cuBLAS GEMMs in the reference.
"""
from __future__ import annotations

import weakref

import torch

from . import _native

_split_cache: "dict[int, tuple]" = {}


def split_weight(weight: torch.Tensor):
    """(w_hi, w_lo) TF32 head / remainder of an (N, K) fp32 weight, cached until the weight is modified or freed."""
    key = id(weight)
    hit = _split_cache.get(key)
    if hit is not None and hit[0]() is weight and hit[1] == weight._version and hit[2] == weight.data_ptr():
        return hit[3], hit[4]
    if not weight.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    w = weight.detach()
    if w.dtype != torch.float32:
        raise TypeError("linear_3xtf32: fp32 weights only")
    w = w.contiguous()
    N, K = w.shape
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    with torch.cuda.device(w.device):              # the library launches on the CURRENT device
        _native.check(_native.lib().msda_b200_linear_split_weight_f32(
            w.data_ptr(), N, K, hi.data_ptr(), lo.data_ptr(), torch.cuda.current_stream(w.device).cuda_stream),
            "msda_b200_linear_split_weight_f32")
    ref = weakref.ref(weight, lambda _r, k=key: _split_cache.pop(k, None))
    _split_cache[key] = (ref, weight._version, weight.data_ptr(), hi, lo)
    return hi, lo


def invalidate_caches() -> None:
    """Drop every cached TF32 weight split.  The cache is keyed by the weight's identity, version counter and storage
    pointer; an in-place write through ``param.data`` (EMA, weight-swap hooks, ``_reset_parameters``) bumps none of them, so
    such code must call this (``MSDeformAttn`` and the drop-in layers do it from ``_load_from_state_dict``, ``_apply`` and
    ``train``)."""
    _split_cache.clear()


def linear_3xtf32(x: torch.Tensor, weight: torch.Tensor, bias: "torch.Tensor | None" = None,
                  row_zero: "torch.Tensor | None" = None, out: "torch.Tensor | None" = None,
                  relu: bool = False) -> torch.Tensor:
    """``F.linear(x, weight, bias)`` for fp32 CUDA tensors on the tcgen05 tensor cores with fp32-grade accuracy.

    x (..., K) with a contiguous last dim (row pitch may exceed K); weight (N, K); returns (..., N).
    row_zero: optional bool/uint8 tensor with one entry per row of x: rows flagged non-zero come out as zeros.
    relu: apply ``max(y, 0)`` in the GEMM epilogue (feed-forward ``activation(linear1(x))``); not with ``row_zero``.
    """
    if not x.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    if x.dtype != torch.float32:
        raise TypeError("linear_3xtf32: fp32 activations only")
    K = x.shape[-1]
    N = weight.shape[0]
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1 or x2.stride(0) % 4 != 0 or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()
    M = x2.shape[0]
    hi, lo = split_weight(weight)
    if out is None:
        out = torch.empty(x.shape[:-1] + (N,), dtype=torch.float32, device=x.device)
    y2 = out.view(-1, N) if out.is_contiguous() else out
    if y2.dim() != 2 or y2.shape[0] != M or y2.stride(1) != 1:
        raise ValueError("linear_3xtf32: out must be (M, N) with unit column stride")
    rz = None
    if row_zero is not None:
        rz = row_zero.reshape(-1)
        if rz.dtype == torch.bool:
            rz = rz.view(torch.uint8)
        rz = rz.contiguous()
        if rz.numel() != M:
            raise ValueError("row_zero must have one entry per row of x")
    b = bias.detach().contiguous() if bias is not None else None
    if b is not None and b.data_ptr() % 16 != 0:   # the epilogue reads the bias as float4
        b = b.clone()
    if relu and rz is not None:
        raise ValueError("linear_3xtf32: relu and row_zero cannot be combined")
    with torch.cuda.device(x.device):              # sm count, kernel attributes and the launch follow the current device
        stream = torch.cuda.current_stream(x.device).cuda_stream
        if relu:
            _native.check(_native.lib().msda_b200_linear_relu_f32(
                x2.data_ptr(), x2.stride(0), hi.data_ptr(), lo.data_ptr(), b.data_ptr() if b is not None else None,
                M, N, K, y2.data_ptr(), y2.stride(0), stream), "msda_b200_linear_relu_f32")
        else:
            _native.check(_native.lib().msda_b200_linear_f32(
                x2.data_ptr(), x2.stride(0), hi.data_ptr(), lo.data_ptr(), b.data_ptr() if b is not None else None,
                rz.data_ptr() if rz is not None else None, M, N, K, y2.data_ptr(), y2.stride(0), stream),
                "msda_b200_linear_f32")
    return out


class TensorCoreLinear(torch.nn.Linear):
    """``nn.Linear`` whose CUDA fp32 inference forward runs on the 3xTF32 tcgen05 GEMM.  Same parameters, same state dict,
    same result to fp32 rounding (<= 1e-5 of max|y| against float64, tests/test_proj_gemm_gpu.py); everything the kernel
    does not take -- CPU tensors, other dtypes, ``out_features % 32`` or ``in_features % 16`` non-zero (the 256 -> 1 / 2 /
    8 heads), a forward that needs gradients -- is ``nn.Linear`` itself."""

    def forward(self, x):
        if (x.is_cuda and x.dtype == torch.float32 and self.weight.dtype == torch.float32 and x.numel() > 0
                and self.out_features % 32 == 0 and self.in_features % 16 == 0 and self.out_features <= 1024
                and not (torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad))):
            return linear_3xtf32(x, self.weight, self.bias)
        return torch.nn.functional.linear(x, self.weight, self.bias)


def use_tensor_core_linears(module: torch.nn.Module) -> int:
    """Re-class every plain ``nn.Linear`` under ``module`` to ``TensorCoreLinear`` in place (parameters untouched) and
    return how many were switched.  Meant for the spotter's frozen detection transformer and rescoring head -- the
    proposal MLPs over all S encoder tokens (deformable_transformer.py:137,182-183), the decoder's reference-point and
    control-point MLPs (:476-486) and the prediction heads (detection_transformer_wobackbone.py:199-216) -- which the
    reference runs as fp32 SIMT cuBLAS GEMMs."""
    n = 0
    for m in module.modules():
        if type(m) is torch.nn.Linear:
            m.__class__ = TensorCoreLinear
            n += 1
    return n
