"""`DeformableCompositeTransformerDecoderLayer` -- drop-in for third_party/adet/layers/deformable_transformer.py:326-427.

The point-query decoder layer of DeepSolo: self-attention among the 25 points of a proposal (``attn_intra``), among the
proposals at the same point index (``attn_inter``), multi-scale deformable cross-attention into the encoder memory
(``attn_cross`` -- the B200 ``MSDeformAttn``), and a feed-forward block.  Same constructor, sub-module and parameter
names (reference state dicts load with ``strict=True``) and the same ``forward`` signature.  The two multi-head
self-attentions keep their ``nn.MultiheadAttention`` parameters; at inference on fp32 CUDA tensors they run as tcgen05
projection GEMMs around one small attention kernel (``small_mha.py``: no transposed copies, no discarded attention
weights), the residual + LayerNorm pairs as one kernel (``norm.add_layernorm``) and the feed-forward GEMMs on the tensor
cores (``projections.linear_3xtf32``), exactly as in ``encoder_layer.py``.
"""
from __future__ import annotations

import torch
from torch import nn

from .encoder_layer import _get_activation_fn
from .ms_deform_attn import CacheInvalidationMixin, MSDeformAttn
from .norm import add_layernorm, add_layernorm_supported
from .projections import linear_3xtf32
from . import small_mha


class DeformableCompositeTransformerDecoderLayer(CacheInvalidationMixin, nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        # self attention (intra: over the points of one proposal)
        self.attn_intra = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.norm_intra = nn.LayerNorm(d_model)
        self.dropout_intra = nn.Dropout(dropout)
        # self attention (inter: over the proposals at one point index)
        self.attn_inter = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout_inter = nn.Dropout(dropout)
        self.norm_inter = nn.LayerNorm(d_model)
        # cross attention
        self.attn_cross = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout_cross = nn.Dropout(dropout)
        self.norm_cross = nn.LayerNorm(d_model)
        # ffn
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _get_activation_fn(activation)
        self._relu = activation == "relu"
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)
        self.tensor_core_ffn = True
        self.fused_add_norm = True
        self.fused_self_attention = True      # attn_intra / attn_inter on this library's kernels at inference
        self._packed_intra = small_mha.PackedProjection()

    def invalidate_caches(self):
        super().invalidate_caches()
        if getattr(self, "_packed_intra", None) is not None:
            self._packed_intra.key = None

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def _inference_fast_path(self, x, *params) -> bool:
        if not (x.is_cuda and x.dtype == torch.float32):
            return False
        if torch.is_grad_enabled() and (x.requires_grad or any(p is not None and p.requires_grad for p in params)):
            return False
        return True

    def _add_norm(self, x, y, dropout, norm):
        """``norm(x + dropout(y))``; one kernel when nothing needs a gradient and dropout is inactive."""
        if (self.fused_add_norm and not (self.training and dropout.p > 0) and add_layernorm_supported(x, norm)
                and self._inference_fast_path(x, norm.weight) and not (torch.is_grad_enabled() and y.requires_grad)):
            return add_layernorm(x, y, norm)
        return norm(x + dropout(y))

    def forward_ffn(self, tgt):
        d, f = self.linear1.in_features, self.linear1.out_features
        if (self.tensor_core_ffn and self._relu and not (self.training and (self.dropout3.p > 0 or self.dropout4.p > 0))
                and self._inference_fast_path(tgt, self.linear1.weight, self.linear2.weight)
                and d % 32 == 0 and f % 32 == 0 and d <= 1024 and f <= 1024):
            tgt2 = linear_3xtf32(linear_3xtf32(tgt, self.linear1.weight, self.linear1.bias, relu=True),
                                 self.linear2.weight, self.linear2.bias)
        else:
            tgt2 = self.linear2(self.dropout3(self.activation(self.linear1(tgt))))
        return self._add_norm(tgt, tgt2, self.dropout4, self.norm3)

    accepts_spatial_shapes_list = True      # transformer_dropin passes the Python shape list down to MSDeformAttn

    def forward(self, tgt, query_pos, reference_points, src, src_spatial_shapes, level_start_index, src_padding_mask=None,
                spatial_shapes_list=None):
        # tgt, query_pos: (bs, n_q, n_pts, d_model)
        bs, n_q, n_pts, dim = tgt.shape
        fast = (self.fused_self_attention and self._inference_fast_path(tgt, self.attn_intra.in_proj_weight)
                and tgt.is_contiguous() and small_mha.supported(self.attn_intra, tgt, n_pts)
                and small_mha.supported(self.attn_inter, tgt, n_q))
        if fast:
            # tokens stay in their (batch, proposal, point) row order; the attention kernel addresses sequences by strides
            T = bs * n_q * n_pts
            tok = tgt.view(T, dim)
            qk = self.with_pos_embed(tgt, query_pos).reshape(T, dim)
            intra = small_mha.self_attention(self.attn_intra, qk, tok, bs * n_q, n_pts, n_pts, 1, self._packed_intra)
            tgt = self._add_norm(tgt, intra.view(bs, n_q, n_pts, dim), self.dropout_intra, self.norm_intra)
            # inter: sequence = the n_q proposals at one point index p of one image: rows (b * n_q + q) * n_pts + p
            if bs == 1:
                inter = small_mha.self_attention(self.attn_inter, tgt.view(T, dim), None, n_pts, n_q, 1, n_pts,
                                                 self._packed_intra)
            else:
                inter = torch.cat([small_mha.self_attention(self.attn_inter, tgt[b].reshape(n_q * n_pts, dim), None, n_pts, n_q,
                                                            1, n_pts, self._packed_intra) for b in range(bs)], 0)
            tgt_inter = self._add_norm(tgt, inter.view(bs, n_q, n_pts, dim), self.dropout_inter, self.norm_inter)
        else:
            # ---- intra: sequences of n_pts points, one per (batch, proposal); nn.MultiheadAttention wants (L, B, E)
            qk = self.with_pos_embed(tgt, query_pos).reshape(bs * n_q, n_pts, dim).transpose(0, 1)
            val = tgt.reshape(bs * n_q, n_pts, dim).transpose(0, 1)
            intra = self.attn_intra(qk, qk, val)[0].transpose(0, 1).reshape(bs, n_q, n_pts, dim)
            tgt = self._add_norm(tgt, intra, self.dropout_intra, self.norm_intra)
            # ---- inter: sequences of n_q proposals, one per (batch, point index)
            t = tgt.transpose(1, 2)                                   # (bs, n_pts, n_q, dim)
            seq = t.reshape(bs * n_pts, n_q, dim).transpose(0, 1)
            inter = self.attn_inter(seq, seq, seq)[0].transpose(0, 1).reshape(bs, n_pts, n_q, dim)
            t = self._add_norm(t.contiguous(), inter, self.dropout_inter, self.norm_inter)
            tgt_inter = t.transpose(1, 2)                             # back to (bs, n_q, n_pts, dim)
        # ---- cross attention into the encoder memory
        if reference_points.dim() == 4:                           # one reference point per proposal: shared by its points
            ref = reference_points[:, :, None, :, :].repeat(1, 1, n_pts, 1, 1)
        else:
            assert reference_points.shape[2] == n_pts
            ref = reference_points
        cross = self.attn_cross(self.with_pos_embed(tgt_inter, query_pos).flatten(1, 2), ref.flatten(1, 2), src,
                                src_spatial_shapes, level_start_index, src_padding_mask,
                                spatial_shapes_list=spatial_shapes_list).reshape(bs, n_q, n_pts, dim)
        tgt = self._add_norm(tgt_inter.contiguous(), cross, self.dropout_cross, self.norm_cross)
        # ---- ffn
        return self.forward_ffn(tgt)
