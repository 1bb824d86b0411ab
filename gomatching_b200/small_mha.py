"""nn.MultiheadAttention for the short sequences of DeepSolo's point-query decoder, on this library's kernels.

Reference: ``attn_intra`` / ``attn_inter`` of ``DeformableCompositeTransformerDecoderLayer``
(third_party/adet/layers/deformable_transformer.py:386-404) are ``nn.MultiheadAttention(d_model, n_heads)`` called as
``attn(q, k, v)[0]`` on (L, B, E) tensors -- torch's functional path with ``need_weights=True``: packed input
projection, ``q * d**-0.5``, ``bmm``, softmax, ``bmm``, head-averaged weights (discarded), output projection.
Here: the input projection is one or two 3xTF32 tcgen05 GEMMs (``projections.linear_3xtf32``) over the tokens in the
order they already have in memory, the attention core is ``msda_b200_small_mha_f32`` (one kernel, strided row
addressing instead of transposed copies), the output projection another GEMM.  Inference only (no autograd), fp32,
head_dim 32, L <= 128, no masks, no bias_k / add_zero_attn -- anything else returns ``None`` and the caller keeps
``nn.MultiheadAttention``.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _native
from .projections import linear_3xtf32


def supported(mha: nn.MultiheadAttention, x: torch.Tensor, L: int) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and mha._qkv_same_embed_dim and mha.in_proj_weight is not None
            and mha.in_proj_weight.dtype == torch.float32 and mha.bias_k is None and mha.bias_v is None
            and not mha.add_zero_attn and mha.head_dim == 32 and L <= 128 and mha.embed_dim % 32 == 0
            and mha.embed_dim <= 1024 and not (mha.training and mha.dropout > 0)
            and not (torch.is_grad_enabled() and (x.requires_grad or mha.in_proj_weight.requires_grad)))


class PackedProjection:
    """``[W_q; W_k]`` / ``W_v`` slices of ``in_proj_weight`` as stand-alone tensors (the TF32-split cache is keyed by
    tensor identity, and a fresh slice per call would miss it every time); rebuilt when the parameter changes."""

    def __init__(self):
        self.key = None
        self.w_qk = self.b_qk = self.w_v = self.b_v = None

    def get(self, mha: nn.MultiheadAttention):
        w, b = mha.in_proj_weight, mha.in_proj_bias
        key = (w._version, w.data_ptr(), None if b is None else (b._version, b.data_ptr()), w.device)
        if key != self.key:
            E = mha.embed_dim
            with torch.no_grad():
                self.w_qk = w[:2 * E].contiguous()
                self.w_v = w[2 * E:].contiguous()
                self.b_qk = None if b is None else b[:2 * E].contiguous()
                self.b_v = None if b is None else b[2 * E:].contiguous()
            self.key = key
        return self


def self_attention(mha: nn.MultiheadAttention, qk_tokens: torch.Tensor, v_tokens: Optional[torch.Tensor], B: int, L: int,
                   batch_stride: int, seq_stride: int, packed: PackedProjection) -> torch.Tensor:
    """Tokens are rows of a contiguous (T, E) matrix, T = B * L; sequence b, position i is row
    ``b * batch_stride + i * seq_stride``.  ``qk_tokens`` feeds the query and key projections; ``v_tokens`` the value
    projection (None: the same rows, one packed GEMM).  Returns the attention output after ``out_proj`` as (T, E) in the
    same row order."""
    T, E = qk_tokens.shape
    H = mha.num_heads
    if v_tokens is None:
        qkv = linear_3xtf32(qk_tokens, mha.in_proj_weight, mha.in_proj_bias)           # (T, 3E)
        q, k, v, ld = qkv, qkv[:, E:], qkv[:, 2 * E:], 3 * E
    else:
        pk = packed.get(mha)
        qkv = torch.empty((T, 3 * E), dtype=torch.float32, device=qk_tokens.device)
        linear_3xtf32(qk_tokens, pk.w_qk, pk.b_qk, out=qkv[:, :2 * E])
        linear_3xtf32(v_tokens, pk.w_v, pk.b_v, out=qkv[:, 2 * E:])
        q, k, v, ld = qkv, qkv[:, E:], qkv[:, 2 * E:], 3 * E
    ctx = torch.empty((T, E), dtype=torch.float32, device=qk_tokens.device)
    with torch.cuda.device(qk_tokens.device):
        rc = _native.lib().msda_b200_small_mha_f32(q.data_ptr(), k.data_ptr(), v.data_ptr(), ld, ctx.data_ptr(), E, B, L, H,
                                                   E // H, batch_stride, seq_stride,
                                                   torch.cuda.current_stream().cuda_stream)
    _native.check(rc, "msda_b200_small_mha_f32")
    return linear_3xtf32(ctx, mha.out_proj.weight, mha.out_proj.bias)
