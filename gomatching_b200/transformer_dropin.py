"""Synchronisation-free drop-ins for the reference's ``DeformableTransformer`` and its encoder stack.

Reference (third_party/adet/layers/deformable_transformer.py):
    DeformableTransformer.forward                    :150-215   flattens the levels, builds ``spatial_shapes`` as a DEVICE
                                                                 tensor from a Python list (:157-169), runs encoder,
                                                                 proposal generation, top-k, decoder
    DeformableTransformer.gen_encoder_output_proposals :108-139 iterates ``for (H_, W_) in spatial_shapes`` over that
                                                                 DEVICE tensor: every ``H_`` is a 0-d CUDA tensor, so the
                                                                 slicing / ``view`` / ``linspace`` calls each read it
                                                                 back -- about 40 host synchronisations per frame
    DeformableTransformerEncoder.get_reference_points :287-300  same loop, same synchronisations
The arithmetic is untouched here; only the loop bounds come from the Python ints the caller already had
(``src.shape[-2:]``).  The tensor operations, their order and their operands are the reference's, so results are
bit-identical (tests/test_transformer_dropin.py), the forward no longer touches the host, and it can be captured in a
CUDA graph (``gomatching_b200.video.spotter_graph``).  The Python shape list is also handed to every ``MSDeformAttn``
call, which gives the TMA window kernel its level geometry without a device->host read.

The classes are made by SUBCLASSING the reference's own classes at install time (``install_into_adet(level=
"transformer")``): constructors, parameters, initialisers and state-dict keys are inherited, not restated.
"""
from __future__ import annotations

import torch

from . import decoder_glue

__all__ = ["make_dropin_classes"]


def _call_layer(layer, *args, shapes_list):
    """Layers of this package accept ``spatial_shapes_list``; the reference's own layers (levels "op" / "module") do not."""
    if getattr(layer, "accepts_spatial_shapes_list", False):
        return layer(*args, spatial_shapes_list=shapes_list)
    return layer(*args)


def make_dropin_classes(dt_module):
    """dt_module: the imported ``adet.layers.deformable_transformer``.  Returns (DeformableTransformer,
    DeformableTransformerEncoder) subclasses with host-free forwards."""
    RefTransformer = dt_module.DeformableTransformer
    RefEncoder = dt_module.DeformableTransformerEncoder
    upcast = dt_module.upcast

    class DeformableTransformerEncoder(RefEncoder):
        """deformable_transformer.py:280-323 with Python-int loop bounds."""

        @staticmethod
        def get_reference_points(spatial_shapes, valid_ratios, device):
            # :287-300; ``spatial_shapes`` is a list of (H, W) Python ints here
            reference_points_list = []
            for lvl, (H_, W_) in enumerate(spatial_shapes):
                H_, W_ = int(H_), int(W_)
                ref_y, ref_x = torch.meshgrid(torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32, device=device),
                                              torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32, device=device),
                                              indexing="ij")
                ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H_)
                ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W_)
                reference_points_list.append(torch.stack((ref_x, ref_y), -1))
            reference_points = torch.cat(reference_points_list, 1)
            return reference_points[:, :, None] * valid_ratios[:, None]

        def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None,
                    spatial_shapes_list=None):
            if spatial_shapes_list is None:                   # called the reference's way: one read-back, then host-free
                spatial_shapes_list = [tuple(int(v) for v in hw) for hw in spatial_shapes.tolist()]
            output = src
            reference_points = self.get_reference_points(spatial_shapes_list, valid_ratios, device=src.device)
            for layer in self.layers:
                output = _call_layer(layer, output, pos, reference_points, spatial_shapes, level_start_index, padding_mask,
                                     shapes_list=spatial_shapes_list)
            return output

    class DeformableTransformer(RefTransformer):
        """deformable_transformer.py:22-215: constructor inherited; forward / proposal generation host-free."""

        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            # the reference builds its encoder container inside __init__ (:57); re-class it in place (same layers)
            self.encoder.__class__ = DeformableTransformerEncoder
            self._shape_cache = {}

        def _shape_tensors(self, shapes_list, device):
            """spatial_shapes / level_start_index as device tensors, built once per (pyramid, device): :169-170."""
            key = (tuple(shapes_list), str(device))
            hit = self._shape_cache.get(key)
            if hit is None:
                if torch.device(device).type == "cuda" and torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("DeformableTransformer: run one forward at this frame size before capturing a CUDA graph")
                if len(self._shape_cache) > 32:
                    self._shape_cache.clear()
                spatial_shapes = torch.as_tensor(shapes_list, dtype=torch.long, device=device)
                level_start_index = torch.cat((spatial_shapes.new_zeros((1,)), spatial_shapes.prod(1).cumsum(0)[:-1]))
                hit = self._shape_cache[key] = (spatial_shapes, level_start_index)
            return hit

        def gen_encoder_output_proposals(self, memory, memory_padding_mask, spatial_shapes):
            # :108-139; ``spatial_shapes``: list of Python (H, W), or the device tensor (read back once)
            if isinstance(spatial_shapes, torch.Tensor):
                spatial_shapes = [tuple(int(v) for v in hw) for hw in spatial_shapes.tolist()]
            N_, S_, C_ = memory.shape
            proposals = []
            _cur = 0
            for lvl, (H_, W_) in enumerate(spatial_shapes):
                mask_flatten_ = memory_padding_mask[:, _cur:(_cur + H_ * W_)].view(N_, H_, W_, 1)
                valid_H = torch.sum(~mask_flatten_[:, :, 0, 0], 1)
                valid_W = torch.sum(~mask_flatten_[:, 0, :, 0], 1)
                grid_y, grid_x = torch.meshgrid(torch.linspace(0, H_ - 1, H_, dtype=torch.float32, device=memory.device),
                                                torch.linspace(0, W_ - 1, W_, dtype=torch.float32, device=memory.device),
                                                indexing="ij")
                grid = torch.cat([grid_x.unsqueeze(-1), grid_y.unsqueeze(-1)], -1)
                scale = torch.cat([valid_W.unsqueeze(-1), valid_H.unsqueeze(-1)], 1).view(N_, 1, 1, 2)
                grid = (grid.unsqueeze(0).expand(N_, -1, -1, -1) + 0.5) / scale
                proposal = grid.repeat(1, 1, 1, 4)
                proposals.append(proposal.view(N_, -1, 8))
                _cur += H_ * W_
            output_proposals = torch.cat(proposals, 1)
            output_proposals_valid = ((output_proposals > 0.01) & (output_proposals < 0.99)).all(-1, keepdim=True)
            output_proposals = torch.log(output_proposals / (1 - output_proposals))
            output_proposals = output_proposals.masked_fill(memory_padding_mask.unsqueeze(-1), float('inf'))
            output_proposals = output_proposals.masked_fill(~output_proposals_valid, float('inf'))
            output_memory = memory
            output_memory = output_memory.masked_fill(memory_padding_mask.unsqueeze(-1), float(0))
            output_memory = output_memory.masked_fill(~output_proposals_valid, float(0))
            output_memory = self.enc_output_norm(self.enc_output(output_memory))
            return output_memory, output_proposals

        def forward(self, srcs, masks, pos_embeds, query_embed):
            # :150-170
            src_flatten, mask_flatten, lvl_pos_embed_flatten, shapes_list = [], [], [], []
            for lvl, (src, mask, pos_embed) in enumerate(zip(srcs, masks, pos_embeds)):
                bs, c, h, w = src.shape
                shapes_list.append((int(h), int(w)))
                src = src.flatten(2).transpose(1, 2)
                mask = mask.flatten(1)
                pos_embed = pos_embed.flatten(2).transpose(1, 2)
                lvl_pos_embed_flatten.append(pos_embed + self.level_embed[lvl].view(1, 1, -1))
                src_flatten.append(src)
                mask_flatten.append(mask)
            src_flatten = torch.cat(src_flatten, 1)
            mask_flatten = torch.cat(mask_flatten, 1)
            lvl_pos_embed_flatten = torch.cat(lvl_pos_embed_flatten, 1)
            spatial_shapes, level_start_index = self._shape_tensors(shapes_list, src_flatten.device)
            valid_ratios = torch.stack([self.get_valid_ratio(m) for m in masks], 1)

            # :172-179
            memory = self.encoder(src_flatten, spatial_shapes, level_start_index, valid_ratios, lvl_pos_embed_flatten,
                                  mask_flatten, spatial_shapes_list=shapes_list)

            # :181-199
            bs, _, c = memory.shape
            output_memory, output_proposals = self.gen_encoder_output_proposals(memory, mask_flatten, shapes_list)
            enc_outputs_class = self.bezier_class_embed(output_memory)
            enc_outputs_coord_unact = self.bezier_coord_embed(output_memory) + output_proposals
            topk = self.num_proposals
            topk_proposals = torch.topk(enc_outputs_class[..., 0], topk, dim=1)[1]
            topk_coords_unact = torch.gather(enc_outputs_coord_unact, 1, topk_proposals.unsqueeze(-1).repeat(1, 1, 8))
            topk_coords_unact = topk_coords_unact.detach()
            reference_points = topk_coords_unact.sigmoid()
            reference_points = self.init_points_from_bezier_proposals(reference_points)
            init_reference_out = reference_points

            # :201-215
            query_embed = query_embed.unsqueeze(0).expand(bs, -1, -1, -1)
            hs, inter_references = self._decode(query_embed, reference_points, memory, spatial_shapes, level_start_index,
                                                valid_ratios, mask_flatten, shapes_list)
            return hs, init_reference_out, inter_references, enc_outputs_class, enc_outputs_coord_unact

        def init_points_from_bezier_proposals(self, reference_bezier):
            # :99-106; the Bernstein matrix is moved to the device once instead of on every call
            bz = reference_bezier.shape[0]
            pts = reference_bezier.view(bz, self.num_proposals, 4, 2)
            if self.bernstein_matrix.device != pts.device:
                if pts.is_cuda and torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("DeformableTransformer: run one forward before capturing a CUDA graph")
                self.bernstein_matrix = self.bernstein_matrix.to(pts.device)
            return torch.matmul(upcast(self.bernstein_matrix), upcast(pts))

        def _decode(self, tgt, reference_points, src, spatial_shapes, level_start_index, valid_ratios, padding_mask,
                    shapes_list):
            """DeformableCompositeTransformerDecoder.forward (:446-497) with the shape list passed down to the layers;
            the decoder container itself (ref_point_head, ctrl_point_coord, layers) is the reference's object."""
            dec = self.decoder
            if not all(getattr(l, "accepts_spatial_shapes_list", False) for l in dec.layers):
                return dec(tgt, reference_points, src, spatial_shapes, level_start_index, valid_ratios, query_pos=None,
                           src_padding_mask=padding_mask)
            output = tgt
            assert reference_points.shape[-1] == 2
            intermediate, intermediate_reference_points = [], []
            # one-launch glue (decoder_glue.py, bit-identical to the eager functions) on CUDA fp32
            fused = decoder_glue.supported(tgt, reference_points, valid_ratios) and not torch.is_grad_enabled()
            for lid, layer in enumerate(dec.layers):
                reference_points_input = reference_points[:, :, :, None] * valid_ratios[:, None, None]
                if fused:
                    query_pos = decoder_glue.point_pos_embed(reference_points, valid_ratios, dec.d_model, dec.temp)
                else:
                    query_pos = dt_module.gen_point_pos_embed(reference_points_input[:, :, :, 0, :], dec.d_model, dec.temp)
                query_pos = dec.ref_point_head(query_pos)
                output = layer(output, query_pos, reference_points_input, src, spatial_shapes, level_start_index,
                               padding_mask, spatial_shapes_list=shapes_list)
                if dec.ctrl_point_coord is not None:
                    tmp = dec.ctrl_point_coord[lid](output)
                    if fused:
                        reference_points = decoder_glue.refine_points(tmp, reference_points)
                    else:
                        new_reference_points = tmp + dt_module.inverse_sigmoid(reference_points)
                        reference_points = new_reference_points.sigmoid().detach()
                if dec.return_intermediate:
                    intermediate.append(output)
                    intermediate_reference_points.append(reference_points)
            if dec.return_intermediate:
                return torch.stack(intermediate), torch.stack(intermediate_reference_points)
            return output, reference_points

    DeformableTransformer.__qualname__ = "DeformableTransformer"
    DeformableTransformerEncoder.__qualname__ = "DeformableTransformerEncoder"
    return DeformableTransformer, DeformableTransformerEncoder
