"""`DeformableTransformerEncoderLayer` -- drop-in for third_party/adet/layers/deformable_transformer.py:218-278.

The caller either side of the operator: self-attention through the B200 `MSDeformAttn`, residual
+ LayerNorm, and the feed-forward block ``linear2(dropout(relu(linear1(src))))`` + residual + LayerNorm.  Same
constructor, same sub-module and parameter names (``self_attn``, ``norm1``, ``linear1``, ``linear2``, ``norm2``: DeepSolo
checkpoints load unchanged), same ``forward`` / ``forward_ffn`` signatures.

At inference on fp32 CUDA tensors the two feed-forward GEMMs run on the tcgen05 tensor cores as 3xTF32 products with
fp32-grade accuracy (``projections.linear_3xtf32``, ReLU fused into the first epilogue); with autograd, dropout
active, other dtypes or ``tensor_core_ffn = False`` the eager reference sequence runs.  The feed-forward block is
where an encoder layer spends its time once the sampler is fast: 161 GFLOP per 720p frame and layer against 0.8 for the
sampling itself.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .ms_deform_attn import CacheInvalidationMixin, MSDeformAttn
from .norm import add_layernorm, add_layernorm_supported
from .projections import linear_3xtf32


def _get_activation_fn(activation):
    """deformable_transformer.py:512-520"""
    if activation == "relu":
        return F.relu
    if activation == "gelu":
        return F.gelu
    if activation == "glu":
        return F.glu
    raise RuntimeError(F"activation should be relu/gelu, not {activation}.")


class DeformableTransformerEncoderLayer(CacheInvalidationMixin, nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        # self attention
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        # ffn
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _get_activation_fn(activation)
        self._relu = activation == "relu"
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.tensor_core_ffn = True
        self.fused_add_norm = True     # inference: residual add + LayerNorm as one kernel (norm.add_layernorm)

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def _ffn_on_tensor_cores(self, src) -> bool:
        if not (self.tensor_core_ffn and self._relu and src.is_cuda and src.dtype == torch.float32):
            return False
        if torch.is_grad_enabled() and (src.requires_grad or self.linear1.weight.requires_grad or self.linear2.weight.requires_grad):
            return False
        if self.training and (self.dropout2.p > 0 or self.dropout3.p > 0):
            return False
        d, f = self.linear1.in_features, self.linear1.out_features
        return d % 32 == 0 and f % 32 == 0 and d <= 1024 and f <= 1024 and self.linear1.weight.dtype == torch.float32

    def forward_ffn(self, src):
        if self._ffn_on_tensor_cores(src):
            hidden = linear_3xtf32(src, self.linear1.weight, self.linear1.bias, relu=True)
            src2 = linear_3xtf32(hidden, self.linear2.weight, self.linear2.bias)
        else:
            src2 = self.linear2(self.dropout2(self.activation(self.linear1(src))))
        return self._add_norm(src, src2, self.dropout3, self.norm2)

    def _add_norm(self, src, src2, dropout, norm):
        """``norm(src + dropout(src2))``; one kernel when nothing needs a gradient and dropout is inactive."""
        needs_grad = torch.is_grad_enabled() and (src.requires_grad or src2.requires_grad or
                                                  (norm.weight is not None and norm.weight.requires_grad))
        if (self.fused_add_norm and not needs_grad and not (self.training and dropout.p > 0)
                and add_layernorm_supported(src, norm)):
            return add_layernorm(src, src2, norm)
        return norm(src + dropout(src2))

    accepts_spatial_shapes_list = True      # transformer_dropin passes the Python shape list down to MSDeformAttn

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None,
                spatial_shapes_list=None):
        # self attention
        src2 = self.self_attn(self.with_pos_embed(src, pos), reference_points, src, spatial_shapes, level_start_index,
                              padding_mask, spatial_shapes_list=spatial_shapes_list)
        src = self._add_norm(src, src2, self.dropout1, self.norm1)
        # ffn
        src = self.forward_ffn(src)
        return src
