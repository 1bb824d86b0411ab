"""gomatching_b200 -- B200-native (sm_100a) multi-scale deformable attention for GoMatching / DeepSolo.

Public surface (mirrors third_party/adet/layers/ms_deform_attn.py of the reference):
    MSDeformAttn                 nn.Module drop-in
    MSDeformAttnFunction         autograd Function drop-in (alias _MSDeformAttnFunction)
    ms_deform_attn_forward       adet._C.ms_deform_attn_forward drop-in
    ms_deform_attn_backward      adet._C.ms_deform_attn_backward drop-in
    ms_deform_attn_forward_fused softmax + offsets->locations + sampler in one kernel
    pair_value_bf16, ms_deform_attn_forward[_fused]_paired   the bf16 operator mode on the neighbour-paired value layout
    DeformableTransformerEncoderLayer  encoder layer drop-in (sampler + tensor-core feed-forward block)
    DeformableCompositeTransformerDecoderLayer  point-query decoder layer drop-in
    install_into_adet            monkey-patch the reference's import sites
    use_tensor_core_linears / accelerate_spotter   switch a built model's remaining nn.Linear layers to the tcgen05 GEMM
"""
from .ms_deform_attn_func import (MSDeformAttnFunction, _MSDeformAttnFunction, fused_supported, locations_softmax,
                                  ms_deform_attn_backward, ms_deform_attn_forward, ms_deform_attn_forward_fused,
                                  ms_deform_attn_forward_fused_paired, ms_deform_attn_forward_paired, pair_value_bf16,
                                  sample_index)
from .ms_deform_attn import MSDeformAttn
from .encoder_layer import DeformableTransformerEncoderLayer
from .decoder_layer import DeformableCompositeTransformerDecoderLayer
from .projections import TensorCoreLinear, use_tensor_core_linears

__all__ = [
    "MSDeformAttn", "MSDeformAttnFunction", "_MSDeformAttnFunction", "ms_deform_attn_forward",
    "ms_deform_attn_backward", "ms_deform_attn_forward_fused", "fused_supported", "sample_index",
    "locations_softmax", "install_into_adet", "pair_value_bf16", "ms_deform_attn_forward_paired",
    "ms_deform_attn_forward_fused_paired", "DeformableTransformerEncoderLayer",
    "DeformableCompositeTransformerDecoderLayer", "TensorCoreLinear", "use_tensor_core_linears", "accelerate_spotter",
]


def install_into_adet(level: str = "layers"):
    """Bind this implementation where the reference looks its operator up.

    ``level`` selects how much is swapped (each level includes the previous one):
      "op"      ``adet._C.ms_deform_attn_forward/backward`` (csrc/vision.cpp:52-55) -> the C-ABI kernels.  Bit-identical
                to the reference CUDA kernel, so every downstream tensor -- and every track ID -- is unchanged.
      "module"  ``adet.layers.ms_deform_attn.MSDeformAttn`` and ``adet.layers.deformable_transformer.MSDeformAttn``
                (deformable_transformer.py:16) -> :class:`MSDeformAttn` (fused glue, tensor-core projections;
                within 1e-4 of the reference module).
      "layers"  ``adet.layers.deformable_transformer.DeformableTransformerEncoderLayer`` (:218) and
                ``DeformableCompositeTransformerDecoderLayer`` (:326) -> the drop-in layers (default).
      "transformer"  additionally ``DeformableTransformer`` / ``DeformableTransformerEncoder`` (:22, :280) -> subclasses
                of the reference's own classes whose forwards never touch the host (transformer_dropin.py): same
                arithmetic, bit-identical results, CUDA-graph capturable, and the level geometry reaches the TMA window
                kernel as Python ints.
      "heads"   the same class bindings as "transformer"; the caller then runs :func:`accelerate_spotter` on the built
                model, which moves the spotter's remaining ``nn.Linear`` layers (proposal MLPs over all encoder tokens,
                decoder reference-point / control-point MLPs, prediction and rescoring heads) to the tensor-core GEMM
                (fp32-grade, not bit-identical to cuBLAS).
    Call after ``adet`` is importable and BEFORE the model is constructed; modules that are not imported yet are
    skipped."""
    import sys
    import types

    if level not in ("op", "module", "layers", "transformer", "heads"):
        raise ValueError("level must be 'op', 'module', 'layers', 'transformer' or 'heads', got %r" % (level,))
    if level == "heads":
        level = "transformer"
    c = sys.modules.get("adet._C")
    if c is None:
        c = types.ModuleType("adet._C")
        sys.modules["adet._C"] = c
        if "adet" in sys.modules:
            sys.modules["adet"]._C = c
    c.ms_deform_attn_forward = ms_deform_attn_forward
    c.ms_deform_attn_backward = ms_deform_attn_backward
    if level == "op":
        return c
    for name in ("adet.layers.ms_deform_attn", "adet.layers.deformable_transformer", "adet.layers"):
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "MSDeformAttn"):
            mod.MSDeformAttn = MSDeformAttn
        if level not in ("layers", "transformer"):
            continue
        if mod is not None and hasattr(mod, "DeformableTransformerEncoderLayer"):
            mod.DeformableTransformerEncoderLayer = DeformableTransformerEncoderLayer
        if mod is not None and hasattr(mod, "DeformableCompositeTransformerDecoderLayer"):
            mod.DeformableCompositeTransformerDecoderLayer = DeformableCompositeTransformerDecoderLayer
    if level == "transformer":
        from .transformer_dropin import make_dropin_classes
        dt = sys.modules.get("adet.layers.deformable_transformer")
        if dt is not None:
            base = getattr(dt, "_msda_b200_reference_classes", None)
            if base is None:                              # remember the reference's own classes: installs are repeatable
                base = dt._msda_b200_reference_classes = (dt.DeformableTransformer, dt.DeformableTransformerEncoder)
            saved = dt.DeformableTransformer, dt.DeformableTransformerEncoder
            dt.DeformableTransformer, dt.DeformableTransformerEncoder = base
            try:
                new_t, new_e = make_dropin_classes(dt)
            finally:
                dt.DeformableTransformer, dt.DeformableTransformerEncoder = saved
            dt.DeformableTransformer, dt.DeformableTransformerEncoder = new_t, new_e
            for name in ("adet.modeling.model.detection_transformer_wobackbone", "adet.modeling.model.detection_transformer"):
                mod = sys.modules.get(name)
                if mod is not None and hasattr(mod, "DeformableTransformer"):
                    mod.DeformableTransformer = new_t
    return c


def accelerate_spotter(model) -> int:
    """Level "heads": re-class the plain ``nn.Linear`` layers of a built GoMatching model's frozen spotter --
    ``model.detection_transformer`` (gom_lstmatcher.py:53) and, if present, ``model.roi_heads.rescoring_head`` (:60-66) --
    to :class:`TensorCoreLinear`.  The tracker's association head is left alone: it runs verbatim.  Returns the number of
    layers switched."""
    n = use_tensor_core_linears(model.detection_transformer)
    head = getattr(getattr(model, "roi_heads", None), "rescoring_head", None)
    if head is not None and getattr(model, "with_rescore", True):
        n += use_tensor_core_linears(head)
    return n
